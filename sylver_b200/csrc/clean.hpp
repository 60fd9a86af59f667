// Matrix checking / cleaning for spldlt_analyse(check = true): restatement of SPRAL's
// clean_cscl_oop for a real symmetric indefinite matrix with a conversion map
// (spral/src/matrix_util.f90:1024-1398 with lmap/map present; heap sort :2729-2777,2840ff;
// apply_conversion_map :2559-2606), as called by the reference's analyse and factorize
// (src/spldlt_analyse_mod.F90:707-739, src/spldlt_factorize_mod.F90:672-679).
#pragma once
#include <vector>

namespace sylver_b200 {

struct CleanMatrix {
   int flag = 0;                 // 0, warnings 1..5 (same numbering as SYLVER_WARNING_*), errors < 0
   int noor = 0;                 // out-of-range entries dropped (entries above the diagonal count)
   int ndup = 0;                 // as the reference reports it (each duplicate is counted twice
                                 // when a map is requested: once when found, once when the list
                                 // of duplicates is appended to the map)
   std::vector<long> ptr;        // n + 1, 1-based
   std::vector<int> row;         // cleaned lower triangle, rows increasing within a column
   std::vector<long> map;        // map[0:ne] = source entry (1-based) of every cleaned entry,
   long lmap = 0;                // then (dest, src) pairs of the duplicates to add; lmap = total length
};

// matrix_util error codes that can come back: -1 allocation, -5 ptr(1) < 1, -6 ptr not
// monotone, -10 a non-empty column with all its entries out of range.
int clean_cscl_oop_sym_indef(int n, const long* ptr_in, const int* row_in, CleanMatrix& out);

// val_out[0:ne] from the caller's val through the map (duplicates summed in map order)
void apply_conversion_map(const CleanMatrix& cm, const double* val, double* val_out);

}  // namespace sylver_b200
