// Matrix scalings computed at factorize (host code): restatements of SPRAL's scaling module
// (spral/src/scaling.f90), which the reference calls from spldlt_factorize
// (src/spldlt_factorize_mod.F90:727-835).
#pragma once
#include <vector>

namespace sylver_b200 {

// options%scaling >= 4: symmetric infinity-norm equilibration (scaling.f90:480-521).
// Returns the iteration count the reference reports.
int equilib_scale_sym(int n, const long* ptr, const int* row, const double* val, double* scaling);

// options%scaling == 2: matching-based scaling by the auction algorithm
// (auction_scale_sym, scaling.f90:269-309 -> auction_match :1504-1609 -> auction_match_core
// :1351-1489 -> match_postproc :1611-1719), default auction_options (:33-38).
struct AuctionInform {
   int flag = 0;
   int matched = 0;
   int iterations = 0;
   int unmatchable = 0;
};
// Lower triangle CSC, 1-based ptr/row.  match (n ints, may be null): match[i] = column (1-based)
// matched to row i, 0 = unmatched.  Returns inform.flag (0, or -1 on allocation failure).
int auction_scale_sym(int n, const long* ptr, const int* row, const double* val, double* scaling, int* match,
                      AuctionInform* inform);

// options%scaling == 1: matching-based scaling by the Hungarian algorithm (MC64-like):
// hungarian_scale_sym (scaling.f90:134-170) -> hungarian_wrapper (:596-801) ->
// hungarian_match (:938-1194) with hungarian_init_heurisitic (:810-929) and the heap (:1206-1325).
struct HungarianInform {
   int flag = 0;      // 0, 1 = WARNING_SINGULAR, -1 allocation, -2 = ERROR_SINGULAR
   int matched = 0;
};
// match (n ints, may be null): match[i] = column (1-based) matched to row i+1; for a structurally
// singular matrix the unmatched rows hold negative values as in the reference.
int hungarian_scale_sym(int n, const long* ptr, const int* row, const double* val, double* scaling, int* match,
                        bool scale_if_singular, HungarianInform* inform);

// The core assignment solver (scaling.f90:938-1194), all arrays 1-based with a dummy element 0:
// minimum-sum matching of the m x n matrix (ptr, row, val >= 0); iperm[i] = column matched to row
// i (negative values complete the matching of a structurally singular matrix), num = cardinality,
// dualu / dualv the dual variables.  Shared with the matching-based ordering (ordering.cpp).
void hungarian_match(int m, int n, const std::vector<long>& ptr, const std::vector<int>& row,
                     const std::vector<double>& val, std::vector<int>& iperm, int& num, std::vector<double>& dualu,
                     std::vector<double>& dualv);

}  // namespace sylver_b200
