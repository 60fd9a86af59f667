// options.ordering = 1: fill-reducing ordering by METIS nested dissection, as the reference
// computes it (spral/src/metis5_wrapper.f90:110-166 metis_order -> METIS_NodeND with default
// options), through the METIS 5 static library that ships with the CUDA toolkit
// (targets/x86_64-linux/lib/libmetis_static.a, idx_t = int64).  Host code.
#pragma once

namespace sylver_b200 {

// true when the library was built with METIS
bool metis_available();

// Lower triangle CSC, 1-based ptr/row (diagonal entries allowed, no duplicates).
// perm[i] = position (1-based) of variable i+1 in the elimination order; invp = its inverse.
// Returns 0, -1 allocation failure, -2 METIS not available, -99 METIS error.
int metis_order(int n, const long* ptr, const int* row, int* perm, int* invp);

// options.ordering = 2: matching-based ordering (spral/src/match_order.f90:135-629
// match_order_metis): Hungarian matching on the expanded matrix, matched pairs kept adjacent,
// METIS on the compressed graph.  Lower triangle CSC with values, 1-based.  order[i] = 1-based
// position of variable i+1; scale (n doubles) = the symmetric MC64-type scaling that
// options.scaling = 3 reuses at factorize.  pairs (n ints, may be null): partner of each
// variable after the cycle splitting (0-based+1; -1 unmatched-in-pair, -2 unmatched).
// Returns 0, 1 (structurally singular: warning), -1 allocation, -2 no METIS, -99.
int match_order_metis(int n, const long* ptr, const int* row, const double* val, int* order, double* scale,
                      int* pairs);

}  // namespace sylver_b200
