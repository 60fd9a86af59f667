// Host side of the B200 numeric factorization engine: the assembly-tree plan
// (SymbolicTree) and the level-scheduled numeric factorization (NumericTree)
// that replace the reference's SymbolicTree / NumericTree[Posdef] + StarPU
// (src/SymbolicTree.cxx, src/NumericTree.hxx:54-406, src/NumericTreePosdef.hxx:42-353).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "../../include/sylver_b200.h"

namespace sylver_b200 {

struct SymbolicTree {
   int n = 0;
   int nnodes = 0;
   long nent = 0;                       // entries of A mapped into fronts
   long nval = 0;                       // extent of the caller's value array (max src index)
   long num_flops = 0;
   // ---- host copies (0-based) ----
   std::vector<int> nrow, ncol, parent, nchild, level;
   std::vector<int> child_ptr, child_list;   // children in decreasing index order (reference order)
   std::vector<long> rptr;                   // 0-based offsets into rlist
   std::vector<long> aent;                   // entries of A mapped into each front
   std::vector<int> rlist;                   // 1-based row indices (as given)
   std::vector<long> cmapoff;                // per node, offset into cmap
   std::vector<int> cmap;                    // child contribution row -> parent local row (0-based)
   std::vector<int> level_ptr, level_nodes;  // fronts grouped by level (height), ncol descending
   // extend-add fused into the contribution epilogue: per front up to two children whose
   // blocks are gathered (largest first), and parent-contribution-row -> child-row maps
   std::vector<int> fchild;                  // [2*nnodes], -1 = none
   std::vector<long> pinvoff;                // [2*nnodes]
   std::vector<int> pinv;
   int nlevels = 0;
   // ---- the reference structure as given at the seam (before chain coarsening) ----
   int ref_nnodes = 0, ref_maxfront = 0;
   std::vector<int> node_map;                // reference node -> engine (coarse) node
   std::vector<int> ref_top;                 // engine node -> its topmost reference node
   std::vector<long> ref_cmapoff;            // assembly maps of the reference structure
   std::vector<int> ref_cmap;
   // ---- device resident static data ----
   int* d_rlist = nullptr;
   long* d_rptr = nullptr;       // 1-based values as given (nnodes+1)
   long* d_nlist = nullptr;
   int* d_anode = nullptr;
   int* d_nrow = nullptr;
   int* d_ncol = nullptr;
   int* d_parent = nullptr;
   int* d_nchild = nullptr;
   int* d_cmap = nullptr;
   long* d_cmapoff = nullptr;
   int* d_level_nodes = nullptr;
   long* d_nptr = nullptr;       // 1-based values as given (nnodes+1)
   int* d_fchild = nullptr;
   long* d_pinvoff = nullptr;
   int* d_pinv = nullptr;
   bool on_device = false;
   int device = 0;

   ~SymbolicTree();
};

// Builds the plan.  Returns nullptr and sets *flag (<0) on failure.  The device
// copy is made lazily (first numeric factorization) so that symbolic parity can
// be tested without a GPU.
SymbolicTree* symbolic_tree_create(int n, int nnodes, const int* sptr, const int* sparent,
                                   const long* rptr, const int* rlist, const long* nptr,
                                   const long* nlist, int* flag);

struct NumericTree;

// One contribution block (or solve work vector) crossing GPUs after a level completes.
struct Xfer {
   int f;      // front whose generated element travels
   int peer;   // the other rank
};
// Deterministic mapping of fronts to ranks and the resulting per-level exchange lists
// (host only; identical on every rank).
void partition_tree(const SymbolicTree& st, int world, std::vector<int>& owner, std::vector<int>* grp0 = nullptr,
                    std::vector<int>* grpn = nullptr);
void plan_exchanges(const SymbolicTree& st, const std::vector<int>& owner, int rank,
                    std::vector<std::vector<Xfer>>& sends, std::vector<std::vector<Xfer>>& recvs);

NumericTree* numeric_tree_create(bool posdef, SymbolicTree* st, const double* aval, const double* scaling,
                                 const sylver_options_c* options, sylver_inform_c* stats);
// Re-run the factorization with new values on an existing tree (same plan); the options are
// read again on every call like the reference does (u, small, pivot methods, action).
void numeric_tree_refactor(NumericTree* nt, const double* aval, const double* scaling,
                           const sylver_options_c* options, sylver_inform_c* stats);
void numeric_tree_destroy(NumericTree* nt);
// job: 1 fwd, 2 diag, 3 bwd, 4 diag+bwd, 0 all.  x host or device, already permuted.
int numeric_tree_solve(const NumericTree* nt, int job, int nrhs, double* x, int ldx);
void numeric_tree_timings(const NumericTree* nt, double* out4);
// out3: fronts split over a rank group in the whole tree, those this rank is a member of, and
// contribution pieces this rank sends per factorization
void numeric_tree_split_info(const NumericTree* nt, int* out3);
int numeric_plan_split(SymbolicTree* st, int rank, int world, long* out8, int cap, long* pieces);
long numeric_plan_levels(SymbolicTree* st, int rank, int world, long cap, long* out);
bool numeric_tree_posdef(const NumericTree* nt);
// Debug / test access: copy one front's L panel (m x n, ld m) and contribution
// ((m-n)^2, ld m-n) to host buffers (either may be null). Returns 0 or <0.
int numeric_tree_get_front(const NumericTree* nt, int node, int* m, int* n, double* l, double* contrib);

// indefinite fronts: eliminated column count, D^-1 (2n doubles) and the pivot permutation (n ints)
int numeric_tree_get_front_indef(const NumericTree* nt, int node, int* nelim, double* d, int* perm);

int device_count();
// `val` if it is host memory, else a host copy in tmp (nullptr if the copy fails)
const double* values_on_host(const double* val, size_t count, std::vector<double>& tmp);
int numeric_tree_profile(const NumericTree* nt, double* out, int cap);
int numeric_tree_profile_levels(const NumericTree* nt, double* out, int cap);
void set_user_stream(void* stream, bool enable);
long numeric_tree_bytes(const NumericTree* nt, long* factor_bytes, long* contrib_bytes);

}  // namespace sylver_b200
