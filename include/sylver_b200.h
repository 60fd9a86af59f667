/* sylver_b200.h -- C ABI of libsylver_b200.so
 *
 * Drop-in boundary for the numeric-factorization path of NLAFET/SyLVER, rebuilt
 * for NVIDIA B200 (sm_100a).  Three groups of entry points:
 *
 *  (1) the public SyLVER C API (binary compatible with the reference's
 *      include/sylver/sylver.h:19-123: same struct layouts, same symbol names),
 *  (2) the internal Fortran->C++ seam the reference's Fortran driver calls
 *      (src/SymbolicTree.cxx:109-137, src/NumericTree.cxx:26-136,
 *      src/NumericTreePosdef.cxx:19-83): a maintainer who keeps the Fortran
 *      analyse/factorize drivers links these instead of the StarPU engine,
 *  (3) sylver_b200_* helpers (introspection, dense single-front drivers,
 *      micro-benchmarks) used by the tests, bench.py and INTEGRATION.md.
 *
 * Conventions (reference: src/interfaces/C/sylver_ciface.F90:374-631):
 *  - A is symmetric, lower triangle in CSC, `long` column pointers, 1-based
 *    ptr/row/order regardless of options.array_base (the reference ignores it).
 *  - Errors are reported in inform->flag (<0 error, >0 warning); nothing throws.
 *  - No CPU fallback exists: without a CUDA device (sm_100) every numeric entry
 *    point sets inform->flag = SYLVER_ERROR_CUDA_UNKNOWN (-51).
 *  - `val` passed to spldlt_factorize / the numeric-tree seam may be a host or a
 *    device pointer (detected with cudaPointerGetAttributes).
 */
#ifndef SYLVER_B200_H
#define SYLVER_B200_H

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------
 * (1) Public API -- replaces include/sylver/sylver.h
 * ------------------------------------------------------------------------- */

/* Layout of reference sylver_inform_t (include/sylver/sylver.h:19-37). */
typedef struct {
   int flag;
   int matrix_dup;
   int matrix_missing_diag;
   int matrix_outrange;
   int matrix_rank;
   int maxdepth;
   int maxfront;
   int num_delay;
   long num_factor;
   long num_flops;
   int num_neg;
   int num_sup;
   int num_two;
   int stat;
   int cuda_error;
   int cublas_error;
   char unused[80];
} sylver_inform_t;

/* Layout of reference sylver_options_t (include/sylver/sylver.h:39-71);
 * defaults from src/sylver_datatypes_mod.F90:97-198. */
typedef struct {
   int array_base;
   int print_level;
   int unit_diagnostics;
   int unit_error;
   int unit_warning;
   int ordering;              /* 0 = user order; 1 = METIS nested dissection (the reference's
                                 default); 2 = matching-based ordering (Hungarian matching,
                                 matched pairs kept adjacent, METIS on the compressed graph;
                                 needs val; its scaling is reused by scaling = 3).  1 and 2 go
                                 through the METIS 5 static library of the CUDA toolkit when the
                                 library was built with it, else flag -98 */
   int nemin;                 /* 32 */
   bool prune_tree;           /* accepted, ignored: every front runs on the GPU (see the seam notes below) */
   long min_gpu_work;
   int scaling;               /* <=0: none / user supplied in `scale`; 1: Hungarian matching
                                 (MC64-like); 2: auction matching; >=4: norm equilibration
                                 (MC77-like), all computed at factorize; 3: the scaling saved by
                                 analyse with ordering = 2 (flag -15 if there is none) */
   int pivot_method;          /* 2 APP block (default): a-posteriori pass, then TPP on what failed;
                                 1 (APP aggressive) and 3 (TPP): TPP on whole fronts, as the
                                 reference's tree code (serial per front, not for performance) */
   double small;              /* 1e-20 */
   double u;                  /* 0.01 */
   long small_subtree_threshold;
   int nb;                    /* default 256; accepted, not used: the engine's blocking is fixed by its kernels
                               * (128-column Cholesky block columns, APTP outer panels of 256 with 32-column
                               * inner blocks = the reference's nb / ib) */
   int cpu_topology;
   bool action;               /* true: continue on singularity with a warning */
   bool use_gpu;
   double gpu_perf_coeff;
   int failed_pivot_method;   /* 1 TPP on failed columns, 2 pass to parent */
   int scheduler;
} sylver_options_t;

/* error / warning codes, src/sylver_datatypes_mod.F90:13-45 */
enum {
   SYLVER_SUCCESS = 0,
   SYLVER_ERROR_CALL_SEQUENCE = -1,
   SYLVER_ERROR_A_N_OOR = -2,
   SYLVER_ERROR_A_PTR = -3,
   SYLVER_ERROR_A_ALL_OOR = -4,
   SYLVER_ERROR_SINGULAR = -5,
   SYLVER_ERROR_NOT_POS_DEF = -6,
   SYLVER_ERROR_PTR_ROW = -7,
   SYLVER_ERROR_ORDER = -8,
   SYLVER_ERROR_VAL = -9,
   SYLVER_ERROR_X_SIZE = -10,
   SYLVER_ERROR_JOB_OOR = -11,
   SYLVER_ERROR_NOT_LLT = -13,
   SYLVER_ERROR_NOT_LDLT = -14,
   SYLVER_ERROR_NO_SAVED_SCALING = -15,
   SYLVER_ERROR_ALLOCATION = -50,
   SYLVER_ERROR_CUDA_UNKNOWN = -51,
   SYLVER_ERROR_CUBLAS_UNKNOWN = -52,
   SYLVER_ERROR_UNIMPLEMENTED = -98,
   SYLVER_ERROR_UNKNOWN = -99,
   SYLVER_WARNING_ANAL_SINGULAR = 6,
   SYLVER_WARNING_FACT_SINGULAR = 7,
   SYLVER_WARNING_MATCH_ORD_NO_SCALE = 8
};

/* sylver.h:73  -- ncpu is accepted and ignored; ngpu = number of B200s this
 * process may use (the one-process-per-GPU launcher passes 1). */
void sylver_init(int ncpu, int ngpu);
/* sylver.h:75 */
void sylver_finalize(void);
/* sylver.h:77 */
void sylver_default_options(sylver_options_t *options);
/* sylver.h:79-83 ; options->ordering 0: order supplied (order[i] = 1-based position of variable
 * i+1); 1: computed by METIS, order may be NULL.  order (if given) is overwritten with the
 * final elimination order.  check = true: the matrix is cleaned first
 * (out-of-range entries dropped, duplicates summed: inform->matrix_outrange / matrix_dup, warning
 * flags 1..5, errors -3 / -4 as the reference); spldlt_factorize then maps the caller's val
 * through the saved conversion map.  check = false: the matrix must be a clean lower triangle. */
void spldlt_analyse(int n, int *order, long const *ptr, int const *row,
                    double const *val, void **akeep, bool check,
                    sylver_options_t const *options, sylver_inform_t *inform);
/* sylver.h:95-98.  val: host or device pointer.  scale (n doubles, original order, may be
 * NULL): read when options->scaling <= 0 (user scaling), written when options->scaling is 1, 2
 * or >= 4 (the scaling computed here, spldlt_factorize_mod.F90:738-795,804-831; needs ptr/row). */
void spldlt_factorize(bool posdef, long const *ptr, int const *row,
                      double const *val, double *scale, void *akeep, void **fkeep,
                      sylver_options_t const *options, sylver_inform_t *inform);
/* sylver.h:100-102 ; job: 0 all, 1 fwd, 2 diag, 3 bwd, 4 diag+bwd */
void spldlt_solve(int job, int nrhs, double *x, int ldx, void *akeep, void *fkeep,
                  sylver_options_t const *options, sylver_inform_t *inform);
/* sylver.h:121,123 */
void spldlt_free_akeep(void **akeep);
void spldlt_free_fkeep(void **fkeep);

/* ---------------------------------------------------------------------------
 * (2) Internal seam -- replaces the StarPU engine behind the Fortran driver
 * ------------------------------------------------------------------------- */

/* Interoperable option / inform subsets, src/sylver_ciface.hxx:39-87 and
 * src/sylver_ciface_mod.F90:9-32. */
typedef struct {
   int print_level;
   bool action;
   double small;
   double u;
   double multiplier;
   long small_subtree_threshold;
   int nb;                    /* accepted, not used (see sylver_options_t.nb) */
   int pivot_method;
   int failed_pivot_method;
   int cpu_topology;
} sylver_options_c;

typedef struct {
   int flag;
   int num_delay;
   int num_neg;
   int num_two;
   int num_zero;
   int maxfront;
   int not_first_pass;
   int not_second_pass;
} sylver_inform_c;

/* src/SymbolicTree.cxx:109-137.  All index arrays are 1-based and BORROWED
 * only during the call (the device plan copies what it needs).
 *
 * Pruned subtrees (the reference's options%prune_tree = .true., its default):
 * the reference hands the nodes listed in `subtrees` to SSIDS through the
 * spldlt_factor_subtree_c callback and receives their generated elements in
 * child_contrib (src/tasks/tasks.hxx:459-513, src/kernels/contrib.cxx:14-38).
 * The B200 engine has no delegate: the seam arrays describe EVERY node, so it
 * factorizes the whole tree itself.  nsubtrees / subtrees / small /
 * contrib_dest / exec_loc are therefore accepted and ignored (any of the
 * arrays may be NULL), child_contrib of the numeric-tree constructors is never
 * read or written, the callback is never called, and the spldlt_tree_solve_*
 * entry points cover all nodes -- a Fortran caller must not run its own
 * subtree factor/solve loops around them (INTEGRATION.md shows the guard), or,
 * equivalently, must analyse with prune_tree = .false. (nsubtrees = 0). */
void *spldlt_create_symbolic_tree(void *akeep, int n, int nnodes, int const *sptr,
                                  int const *sparent, long const *rptr,
                                  int const *rlist, long const *nptr,
                                  long const *nlist, int nsubtrees,
                                  int const *subtrees, int const *small,
                                  int const *contrib_dest, int const *exec_loc);
void spldlt_destroy_symbolic_tree(void *symbolic_tree);

/* src/NumericTree.cxx:26-44: factorization happens inside, synchronous. */
void *spldlt_create_numeric_tree_dbl(bool posdef, void *fkeep, void *symbolic_tree,
                                     double *aval, const double *scaling,
                                     void **child_contrib,
                                     sylver_options_c *options,
                                     sylver_inform_c *stats);
/* src/NumericTreePosdef.cxx:19-35 */
void *spldlt_create_numeric_tree_posdef_dbl(void *fkeep, void *symbolic_tree,
                                            double *aval, const double *scaling,
                                            void **child_contrib,
                                            sylver_options_c *options,
                                            sylver_inform_c *stats);
void spldlt_destroy_numeric_tree_dbl(bool posdef, void *tree);
void spldlt_destroy_numeric_tree_posdef_dbl(void *tree);

/* src/NumericTree.cxx:54-136 (x already permuted to elimination order; host or
 * device pointer).  Return a Flag (0 success). */
int spldlt_tree_solve_fwd_dbl(bool posdef, void const *tree, int nrhs, double *x, int ldx);
int spldlt_tree_solve_bwd_dbl(bool posdef, void const *tree, int nrhs, double *x, int ldx);
int spldlt_tree_solve_diag_dbl(bool posdef, void const *tree, int nrhs, double *x, int ldx);
int spldlt_tree_solve_diag_bwd_dbl(bool posdef, void const *tree, int nrhs, double *x, int ldx);
/* src/NumericTreePosdef.cxx:44-83 */
int spldlt_tree_solve_fwd_posdef_dbl(void const *tree, int nrhs, double *x, int ldx);
int spldlt_tree_solve_bwd_posdef_dbl(void const *tree, int nrhs, double *x, int ldx);

/* ---------------------------------------------------------------------------
 * (3) sylver_b200_* helpers
 * ------------------------------------------------------------------------- */

/* Library / device introspection. Returns number of visible CUDA devices
 * (0 when none; never fails). */
int sylver_b200_device_count(void);
const char *sylver_b200_version(void);

/* Symbolic introspection of an akeep (for bit-exact parity tests).  Each
 * pointer receives a borrowed pointer valid until spldlt_free_akeep. */
typedef struct {
   int n;
   int nnodes;
   int const *sptr;      /* nnodes+1 */
   int const *sparent;   /* nnodes   */
   long const *rptr;     /* nnodes+1 */
   int const *rlist;     /* rptr[nnodes]-1 */
   long const *nptr;     /* nnodes+1 */
   long const *nlist;    /* 2*(nptr[nnodes]-1) */
   int const *order;     /* n */
   int const *invp;      /* n */
   long num_factor;
   long num_flops;
} sylver_b200_symbolic_view;
int sylver_b200_akeep_view(void *akeep, sylver_b200_symbolic_view *view);

/* Per-edge assembly maps derived at symbolic-tree creation:
 * for node c (0-based) cmap[cptr[c] .. cptr[c+1]) gives, for each contribution
 * row of c, its 0-based local row index in the parent front.  Borrowed. */
int sylver_b200_symbolic_tree_cmap(void *symbolic_tree, long const **cptr,
                                   int const **cmap);

/* The SymbolicTree (seam object) owned by an akeep; borrowed. */
void *sylver_b200_akeep_tree(void *akeep);

/* Timings of the last factorization on an fkeep/numeric tree (seconds):
 * out[0]=total device time (CUDA events), out[1]=H2D of aval, out[2]=host
 * wall time of the call, out[3]=kernel launches issued. */
int sylver_b200_numeric_tree_timings(void const *tree, double *out4);
void *sylver_b200_fkeep_tree(void *fkeep);
/* Multi-GPU runs: out3[0] = fronts of the tree that are split over a rank group (block-column
 * cyclic, panel broadcasts; SURVEY.md 8e "top of tree"), out3[1] = those this rank is a member
 * of, out3[2] = contribution-block pieces this rank sends per factorization. */
int sylver_b200_numeric_tree_split_info(void const *tree, int *out3);

/* Per-kernel-class device time of the last factorization when the environment variable
 * SYLVER_B200_PROFILE=1 was set at tree creation (un-graphed issue, CUDA events around
 * every launch): out receives triples (ms, launches, algorithmic flops) for the classes
 * scatter, zero, assemble, potrf|pivot, trsm|tpp, update, contrib.  Returns the number of
 * classes (0 when profiling is off). */
int sylver_b200_numeric_tree_profile(void const *tree, double *out, int cap);
/* Same run, per tree level: out[level * 7 + class] = milliseconds.  Returns the number of levels. */
int sylver_b200_numeric_tree_profile_levels(void const *tree, double *out, int cap);
/* Device bytes held by a numeric tree (total; factor and contribution arenas separately). */
long sylver_b200_numeric_tree_bytes(void const *tree, long *factor_bytes, long *contrib_bytes);
/* Issue all work of subsequently created numeric trees on the caller's CUDA stream
 * (cudaStream_t passed as void*); enable = 0 restores private streams. */
void sylver_b200_set_stream(void *cuda_stream, int enable);
/* Copy one front back to the host (tests): L panel m x n (ld m) and contribution block
 * (m-n)^2 (ld m-n); either may be NULL.  For indefinite trees _indef returns the number
 * of eliminated columns, D^-1 (2n doubles) and the pivot permutation (n ints, 1-based). */
int sylver_b200_numeric_tree_get_front(void const *tree, int node, int *m, int *n, double *l, double *contrib);
/* The engine works on a chain-coarsened copy of the assembly tree (a front that is its
 * parent's last child is merged into it when that adds almost no explicit zeros; environment
 * SYLVER_B200_AMALGAMATE=<fraction>, 0 disables, read at analyse).  `node` in the engine-level
 * calls above counts ENGINE fronts.  This returns the number of reference fronts and the
 * engine's structure: *nnodes fronts, rows / fully-summed columns / parent (0-based, *nnodes =
 * virtual root) per engine front, and node_map[reference front] = engine front.  Borrowed. */
int sylver_b200_symbolic_tree_view(void *symbolic_tree, int *nnodes, int const **nrow, int const **ncol,
                                   int const **parent, int const **node_map);
int sylver_b200_numeric_tree_get_front_indef(void const *tree, int node, int *nelim, double *d, int *perm);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink/NVSwitch -------------------------
 * The reference has no distributed layer (single process, StarPU workers, PCIe staging).
 * Here every rank runs the same analyse; spldlt_factorize then factorizes the fronts the
 * deterministic partition (sylver_b200_partition) gives to this rank and moves the
 * contribution blocks of cross-GPU tree edges with ncclSend/ncclRecv.  spldlt_solve works on a
 * replicated right-hand side and returns the full solution on every rank.  Positive definite
 * factorizations additionally split the widest fronts above the partition cut over their rank
 * group (block-column cyclic, panels by ncclBroadcast on ncclCommSplit sub-communicators;
 * environment SYLVER_B200_SPLIT=0 disables, SYLVER_B200_SPLIT_MIN = minimum columns, default
 * 1024).  Indefinite (APTP) factorizations exchange the eliminated counts, contribution blocks
 * and delayed columns per level; their fronts are not split.
 *   rank 0:    sylver_b200_comm_unique_id(id)   (128 bytes; broadcast it by any means)
 *   all ranks: cudaSetDevice(local gpu); sylver_b200_comm_init(rank, world, id)       */
int sylver_b200_comm_unique_id(void *out128);
int sylver_b200_comm_init(int rank, int world, void const *id128);
void sylver_b200_comm_finalize(void);
int sylver_b200_comm_rank(void);
int sylver_b200_comm_world(void);
/* Planning-only communicator (no NCCL object): lets host code inspect what rank `rank` of
 * `world` would do; numeric calls with it fail. */
void sylver_b200_comm_set_virtual(int rank, int world);
/* Thread-per-rank communicator: `world` threads of ONE process, all on the current device,
 * each call this once (same fabric_id) before factorizing its share and
 * sylver_b200_comm_finalize() before exiting.  The multi-rank schedule (tree partition,
 * contribution/delay hand-over, distributed fronts) then runs over device-to-device copies
 * instead of NCCL -- the transport the single-GPU test box uses.  Thread local. */
int sylver_b200_comm_init_local(int rank, int world, int fabric_id);
/* owner[f] (0-based rank) of every front for `world` GPUs; returns the number of fronts.
 * Replaces the reference's prune_tree / find_subtree_partition
 * (src/spldlt_analyse_mod.F90:949-1141,1435-1656): proportional mapping by flops. */
int sylver_b200_partition(void *akeep, int world, int *owner);
/* The contribution blocks rank `rank` sends/receives, as quadruples
 * (level, front, peer, dir: 0 send / 1 recv); returns their number (may exceed cap/4). */
int sylver_b200_plan_exchanges(void *akeep, int rank, int world, int cap, int *out);
/* Host-only (no GPU, no communicator): the positive definite multi-GPU plan of `rank` among
 * `world` ranks, including the fronts split over rank groups (environment SYLVER_B200_SPLIT,
 * SYLVER_B200_SPLIT_MIN as at factorization).  out8 = { split fronts in the tree, split fronts
 * this rank works on, factor arena bytes, contribution arena bytes, panel staging bytes,
 * contribution pieces sent, received, most point-to-point operations in one level }.  pieces
 * receives 6 longs per piece while they fit in cap: level, front (topmost reference node), peer,
 * offset and count (doubles, inside the front's contribution block), direction (0 send, 1
 * receive).  Returns the number of pieces. */
/* The ordering of options->ordering == 1 on its own (host only): METIS_NodeND with default
 * options on the adjacency lists SPRAL's metis_order builds (spral/src/metis5_wrapper.f90).
 * order[i] = 1-based position of variable i+1, invp its inverse.  Returns 0, -1 (allocation),
 * -2 (built without METIS) or -99. */
int sylver_b200_metis_order(int n, long const *ptr, int const *row, int *order, int *invp);
/* The ordering of options->ordering == 2 on its own (host only): SPRAL match_order_metis
 * (spral/src/match_order.f90:135-629).  order[i] = 1-based position of variable i+1; scale = the
 * symmetric matching scaling (what options->scaling == 3 applies); pairs (n ints, may be NULL):
 * the 1-based partner of every variable in a 2x2 pivot, -1 for a 1x1 pivot, -2 for an unmatched
 * variable.  Returns 0, 1 (structurally singular), -1, -2 (built without METIS) or -99. */
int sylver_b200_match_order(int n, long const *ptr, int const *row, double const *val, int *order, double *scale,
                            int *pairs);
int sylver_b200_plan_split(void *akeep, int rank, int world, long *out8, int cap, long *pieces);
/* Host-only: the level-batched launch plan of the positive definite path for `rank` (fronts that
 * are not split).  Per level 4 longs (level, fronts, block-column steps of 128, contribution
 * tiles), then per step 8 longs (fronts that own the block column, panel-solve tiles, trailing
 * update tiles, of which: first tile column / the rest, first two tile columns / the rest, and
 * the stride of the inverse slots).  Tiles are 128 x 128.  Returns the number of longs. */
long sylver_b200_plan_levels(void *akeep, int rank, int world, long cap, long *out);

/* What spldlt_analyse(check = true) does to the matrix, on its own (host only): SPRAL's
 * clean_cscl_oop for a symmetric indefinite matrix (spral/src/matrix_util.f90:1024-1398):
 * entries above the diagonal or outside 1..n are dropped, duplicates are summed, rows are
 * sorted.  ptr_out: n+1 longs; row_out: cap ints; map: 2*cap longs -- map[0:ne] is the 1-based
 * source entry of every cleaned entry, followed by (destination, source) pairs of the
 * duplicates to add.  counts5 = { flag, out-of-range entries, duplicates as the reference
 * counts them, ne, length of map }.  Returns flag: 0, a warning 1..5 (SYLVER_WARNING_IDX_OOR
 * .. MISS_DIAG_OORDUP) or matrix_util's error (-5/-6 bad ptr, -10 a column with only
 * out-of-range entries). */
int sylver_b200_clean_matrix(int n, long const *ptr, int const *row, int cap, long *ptr_out, int *row_out,
                             long *map, long *counts5);

/* Values of the cleaned matrix from the caller's values through that map (what
 * spldlt_factorize does after a checked analyse; apply_conversion_map, matrix_util.f90:2559). */
int sylver_b200_apply_conversion_map(long ne, long lmap, long const *map, double const *val, double *val_out);

/* The scaling spldlt_factorize computes for options->scaling >= 4, on its own (host only):
 * symmetric infinity-norm equilibration of the lower-triangle CSC matrix (1-based ptr/row),
 * SPRAL inf_norm_equilib_sym (spral/src/scaling.f90:480-521).  scaling: n doubles out.
 * Returns the number of iterations the reference would report, or -1 on bad arguments. */
int sylver_b200_equilib_scale(int n, long const *ptr, int const *row, double const *val, double *scaling);
/* The scaling of options->scaling == 2: matching-based scaling by the auction algorithm with
 * SPRAL's default auction_options (auction_scale_sym, spral/src/scaling.f90:269-309,1351-1719).
 * match (n ints, may be NULL): match[i] = 1-based column matched to row i+1, 0 = unmatched.
 * inform4 (may be NULL) = { flag, matched, iterations, unmatchable }.  Returns 0, or -1. */
int sylver_b200_auction_scale(int n, long const *ptr, int const *row, double const *val, double *scaling,
                              int *match, int *inform4);
/* The scaling of options->scaling == 1: matching-based scaling by the Hungarian algorithm
 * (MC64-like; hungarian_scale_sym, spral/src/scaling.f90:134-170,596-1325).  match as above
 * (negative entries complete the matching of a structurally singular matrix, as in the
 * reference).  inform2 (may be NULL) = { flag, matched }; flag 1 = structurally singular and
 * scale_if_singular, -2 = structurally singular and not scale_if_singular.  Returns flag. */
int sylver_b200_hungarian_scale(int n, long const *ptr, int const *row, double const *val, double *scaling,
                                int *match, int scale_if_singular, int *inform2);

/* Dense single front drivers (reference harness shape:
 * tests/testing_factor_node_indef.hxx:44-460, testing_factor_node_posdef.hxx).
 * a: m x n column-major panel (lda >= m), host memory, overwritten by L;
 * d: 2*n (D^-1 in the reference's storage convention), perm: n (in: 1-based
 * labels, out: permuted), contrib: (m-n)^2 (ld m-n) receives the Schur
 * complement.  Returns nelim (>=0) or a negative SYLVER_ERROR_*.
 * ms_out (may be NULL) receives device milliseconds of the factorization. */
int sylver_b200_factor_front_posdef(int m, int n, double *a, int lda,
                                    double *contrib, int nb, float *ms_out);
int sylver_b200_factor_front_indef(int m, int n, int *perm, double *a, int lda,
                                   double *d, double *contrib,
                                   sylver_options_c const *options,
                                   sylver_inform_c *stats, float *ms_out);

/* Micro-benchmarks run on the current device; return achieved TFLOP/s or GB/s
 * (negative on error).  kind 0: register-resident DMMA issue peak,
 * 1: our tiled DMMA SYRK/GEMM update kernel on an n x n x k problem. */
double sylver_b200_bench_dmma(int kind, int n, int k, int iters);
double sylver_b200_bench_copy(long nbytes, int iters);

#ifdef __cplusplus
}
#endif
#endif /* SYLVER_B200_H */
