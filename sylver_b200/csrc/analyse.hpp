// Symbolic analysis (host, C++) for the B200 numeric factorization engine.
//
// Produces exactly the arrays the reference hands across its Fortran->C++ seam
// (spldlt_create_symbolic_tree, /root/reference/src/SymbolicTree.cxx:109-137):
//   sptr, sparent, rptr, rlist, nptr, nlist  -- all 1-based, as in Fortran.
// Algorithms restate SPRAL's basic_analyse pipeline
// (spral/src/core_analyse.f90:38-150) and SyLVER's build_map
// (src/spldlt_analyse_mod.F90:130-232) so the output is bit-identical.
#pragma once
#include <cstdint>
#include <vector>

namespace sylver_b200 {

struct Symbolic {
   int n = 0;
   int nnodes = 0;
   std::vector<int> sptr;      // nnodes+1, 1-based column starts
   std::vector<int> sparent;   // nnodes,   1-based parent (nnodes+1 = virtual root)
   std::vector<long> rptr;     // nnodes+1, 1-based
   std::vector<int> rlist;     // rptr[nnodes]-1 entries, 1-based row indices (pivot order)
   std::vector<long> nptr;     // nnodes+1, 1-based
   std::vector<long> nlist;    // 2*nz: (src, dest) pairs, 1-based
   std::vector<int> order;     // n: order[i] = pivot position of variable i (1-based; 0 = unused)
   std::vector<int> invp;      // n: inverse of order (pivot position -> variable), 1-based values
   long num_factor = 0;
   long num_flops = 0;
   int maxfront = 0;
   int maxdepth = 0;
   int matrix_rank = 0;
   int realn = 0;
};

// Flags follow src/sylver_datatypes_mod.F90:13-45.
enum AnalyseFlag : int {
   ANAL_SUCCESS = 0,
   ANAL_ERROR_A_N_OOR = -2,
   ANAL_ERROR_A_PTR = -3,
   ANAL_ERROR_A_ALL_OOR = -4,
   ANAL_ERROR_ORDER = -8,
   ANAL_ERROR_ALLOCATION = -50,
   ANAL_WARNING_ANAL_SINGULAR = 6
};

// Lower-triangular CSC (1-based ptr/row) -> full symmetric pattern.
// spral/src/ssids/anal.f90:37-81 (expand_pattern).
void expand_pattern(int n, long nz, const long* ptr, const int* row,
                    std::vector<long>& aptr, std::vector<int>& arow);

// Full analysis given a user pivot order (order[i] = position of variable i, 1-based).
// On return sym.order holds the final elimination order.  Returns an AnalyseFlag.
int analyse(int n, const long* ptr, const int* row, const int* user_order,
            int nemin, Symbolic& sym);

}  // namespace sylver_b200
