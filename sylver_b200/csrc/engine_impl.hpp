// Private definitions shared by engine.cu (plan, posdef path, solves) and
// engine_indef.cu (APTP path).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "comm.hpp"
#include "engine.hpp"
#include "kernels.cuh"

namespace sylver_b200 {

#define CU_TRY(expr)                                                                    \
   do {                                                                                 \
      cudaError_t e__ = (expr);                                                         \
      if (e__ != cudaSuccess) {                                                         \
         fprintf(stderr, "sylver_b200: CUDA error %s at %s:%d (%s)\n", cudaGetErrorName(e__), \
                 __FILE__, __LINE__, #expr);                                            \
         throw CudaFailure{(int)e__};                                                   \
      }                                                                                 \
   } while (0)

struct CudaFailure {
   int code;
};

template <typename T>
static T* dev_upload(const T* h, size_t count) {
   T* d = nullptr;
   CU_TRY(cudaMalloc(&d, std::max<size_t>(count, 1) * sizeof(T)));
   if (count) CU_TRY(cudaMemcpy(d, h, count * sizeof(T), cudaMemcpyHostToDevice));
   return d;
}
template <typename T>
static T* dev_upload(const std::vector<T>& v) {
   return dev_upload(v.data(), v.size());
}

static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }
// CTAs per front for k_zero_contrib: enough to saturate HBM on the widest block of the level
static inline int zero_grid_x(int maxk) {
   const long bytes = (long)maxk * maxk * 8;
   return (int)std::min<long>(std::max<long>(bytes / (256 * 16 * 8), 1), 1184);
}

// ===========================================================================
// Contribution arena planning (static: m - n never changes, even with delays)
// ===========================================================================
namespace {
struct SegAlloc {
   // first-fit free list over [0, inf), sizes in doubles
   std::map<long, long> free_;   // offset -> size
   long top = 0;
   long peak = 0;
   long alloc(long sz) {
      for (auto it = free_.begin(); it != free_.end(); ++it) {
         if (it->second >= sz) {
            long off = it->first;
            long rem = it->second - sz;
            free_.erase(it);
            if (rem > 0) free_[off + sz] = rem;
            return off;
         }
      }
      // extend the top (merge with a trailing free segment if adjacent)
      long off = top;
      if (!free_.empty()) {
         auto last = std::prev(free_.end());
         if (last->first + last->second == top) {
            off = last->first;
            free_.erase(last);
         }
      }
      top = off + sz;
      peak = std::max(peak, top);
      return off;
   }
   void release(long off, long sz) {
      auto it = free_.emplace(off, sz).first;
      auto nx = std::next(it);
      if (nx != free_.end() && it->first + it->second == nx->first) {
         it->second += nx->second;
         free_.erase(nx);
      }
      if (it != free_.begin()) {
         auto pv = std::prev(it);
         if (pv->first + pv->second == it->first) {
            pv->second += it->second;
            free_.erase(it);
         }
      }
   }
};
}  // namespace

// ===========================================================================
// NumericTree
// ===========================================================================
struct LevelStep {
   int cnt;             // fronts of the level that own block column `s`
   int wld;             // stride of the inverse slots for this step
   int trsm_tiles, upd_tiles;
   size_t trsm_prefix, upd_prefix;   // offsets into d_prefix
   int updn_tiles = 0, updr_tiles = 0;      // look-ahead split: first tile column / the rest
   size_t updn_prefix = 0, updr_prefix = 0;
   int upd2n_tiles = 0, upd2r_tiles = 0;    // paired (rank-2nb) updates: first two tile columns / the rest
   size_t upd2n_prefix = 0, upd2r_prefix = 0;
};
struct LevelPlan {
   int first, count;               // range in level_nodes
   int max_children;
   int max_contrib = 0;            // largest contribution block order among fronts with children
   std::vector<std::pair<size_t, int>> asm_work;   // per child ordinal: (offset, count) in d_asm_work
   std::vector<LevelStep> steps;
   int contrib_tiles;
   size_t contrib_prefix;
};

// One contiguous piece of a contribution block crossing GPUs (offset relative to coff[f]).
struct Piece {
   int f, peer;
   long off;
   size_t count;
   long seq;      // position of this (piece, destination) pair in the enumeration all ranks share
};
// A front whose block columns are dealt round-robin to the ranks [g0, g0 + P) (top of the
// tree): this rank is member q; panels travel by broadcast on sub-communicator gid.
struct SplitPlan {
   int f = -1, P = 1, q = 0, g0 = 0, gid = -1, slot = 0;
   std::vector<std::pair<size_t, int>> asm_work;
};

enum KClass { KC_SCATTER = 0, KC_ZERO, KC_ASSEMBLE, KC_POTRF, KC_TRSM, KC_UPDATE, KC_CONTRIB, KC_COUNT };


struct NumericTree {
   using Xfer = ::sylver_b200::Xfer;
   SymbolicTree* st = nullptr;
   bool posdef = true;
   sylver_options_c opt{};
   int nb = 128;
   // per-front geometry (host) + device mirrors
   std::vector<int> m, n, ldl, ldc;
   std::vector<long> loff, coff;
   int *d_m = nullptr, *d_n = nullptr, *d_ldl = nullptr, *d_ldc = nullptr;
   long *d_loff = nullptr, *d_coff = nullptr;
   double* d_L = nullptr; size_t L_doubles = 0;
   double* d_C = nullptr; size_t C_doubles = 0;
   double* d_W = nullptr; size_t W_doubles = 0;
   double* d_aval = nullptr; size_t aval_count = 0;
   double* d_scaling = nullptr;
   int* d_fail = nullptr;
   int* d_prefix = nullptr;
   int2* d_asm_work = nullptr;
   std::vector<LevelPlan> levels;
   DevTree T{};
   cudaStream_t stream = nullptr;
   bool own_stream = true;
   cudaStream_t stream2 = nullptr;       // look-ahead: next panel's potrf + solve run beside the update (high priority)
   cudaStream_t stream3 = nullptr;       // APTP: per-panel contribution passes beside the pivoting chain (low priority)
   cudaEvent_t ev_next = nullptr, ev_panel = nullptr;
   cudaGraphExec_t graph = nullptr;
   bool graph_multi = false;             // world > 1: the NCCL launch sequence is captured too
   // profiling (SYLVER_B200_PROFILE=1): per-class device time / launches / algorithmic flops
   bool profile = false;
   bool pair_updates = true;             // SYLVER_B200_PAIR=0: one rank-nb trailing update per block column
   bool potrf_reg = false;               // SYLVER_B200_POTRF_REG=1: register-resident column-at-a-time kernel (A/B runs)
   bool potrf_old = false;               // SYLVER_B200_POTRF_OLD=1: shared-memory k_potrf_inv<128> (A/B runs)
   std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> prof_events;   // class + 256 * level
   int prof_level = 0;                   // tree level being issued (tags prof_events)
   std::vector<double> prof_level_ms;    // [level * KC_COUNT + class] of the last profiled run
   double prof_ms[KC_COUNT] = {0};
   long prof_launches[KC_COUNT] = {0};
   double prof_flops[KC_COUNT] = {0};
   cudaEvent_t ev0 = nullptr, ev1 = nullptr;
   long launches = 0;
   double t_device = 0, t_h2d = 0, t_wall = 0;
   // solve workspace
   double* d_xw = nullptr;        // sum of m doubles
   long* d_xwoff = nullptr;
   std::vector<long> xwoff;
   int* d_child_ptr = nullptr; int* d_child_list = nullptr;
   int* d_solve_bar = nullptr;           // multi-CTA solve: barrier counters (one per front slot of a launch)
   double* d_solve_part = nullptr;       //                  partial dot products [slot][G][128]
   double* d_xrhs = nullptr; size_t xrhs_cap = 0;   // device copy of a host right-hand side (kept between calls)
   // ---- multi-GPU (one process per GPU): fronts owned by this rank, exchanges per level ----
   int rank = 0, world = 1;
   std::vector<int> owner;               // rank owning each front (partition_tree)
   std::vector<int> lvl_ptr, lvl_nodes;  // owned fronts grouped by level (ncol descending)
   int* d_lvl_nodes = nullptr;
   std::vector<std::vector<Xfer>> sends, recvs;   // per level: contribution blocks crossing GPUs
   int* d_owner = nullptr;
   int* d_all_nodes = nullptr;           // st->level_nodes (all fronts), for the solve broadcasts
   int* d_xpack_off = nullptr;
   double* d_xbuf = nullptr; size_t xbuf_cap = 0;
   // ---- fronts split over a rank group (posdef): block-column cyclic, panel broadcasts ----
   std::vector<int> grp0, splitP, splitQ;         // per front; splitP == 1: not split
   int *d_splitP = nullptr, *d_splitQ = nullptr;
   int* d_scat_owner = nullptr;                   // == rank for every front this rank holds a copy of
   std::vector<int> fac_ptr, fac_nodes;           // owned, unsplit fronts per level (the batched path)
   int* d_fac_nodes = nullptr;
   std::vector<std::vector<SplitPlan>> splits;    // per level: split fronts this rank is a member of
   std::vector<std::vector<Piece>> csends, crecvs;   // per level: contribution pieces crossing GPUs
   std::vector<int> split_fronts;                 // fronts of `splits`, in slot order
   int* d_split_fronts = nullptr;
   int* d_zero = nullptr;
   double* d_stage = nullptr; size_t stage_doubles = 0;
   double* d_Wsplit = nullptr;
   // ---- indefinite (APTP) path: dynamic geometry, see engine_indef.cu ----
   struct Chunk { double* ptr; size_t cap, used; };
   std::vector<Chunk> chunks;            // factor arena: L panel + D^-1 + perm per front, bump allocated
   std::vector<int> nelim;               // host copy, per front (valid after its level completed)
   std::vector<long> woff, doff, permoff;
   int* d_ncol0 = nullptr;
   long *d_woff = nullptr, *d_doff = nullptr, *d_permoff = nullptr;
   FrontState* d_state = nullptr;
   int* d_nelim = nullptr;               // per front (device), for the solves
   int* d_stats = nullptr;               // 8 ints, see k_front_stats
   void* d_diag = nullptr;               // DiagScratch per front of the widest level
   size_t diag_cap = 0;
   void* d_lvl = nullptr;                // per-level upload buffer (geometry updates, orders, prefixes)
   void* h_lvl = nullptr;                // pinned mirror
   size_t lvl_cap = 0;
   cudaEvent_t ev_bulk = nullptr;        // bulk part of a panel's trailing update on stream3 (APTP look-ahead)
   cudaEvent_t ev_cpass[2] = {nullptr, nullptr};   // per-panel contribution passes on stream2 (APTP)
   cudaEvent_t ev_lvl = nullptr;         // completion of the last H2D copy out of h_lvl
   bool lvl_busy = false;
   int* d_lvl_out = nullptr;             // nelim of the level's fronts (read back per level)
   int* h_lvl_out = nullptr;
   size_t lvl_out_cap = 0;
   size_t Wscratch_cap = 0;
   int num_delay = 0, maxfront = 0;
};

namespace {
struct ProfScope {
   NumericTree* nt; int cls; cudaEvent_t a = nullptr, b = nullptr;
   cudaStream_t st;
   ProfScope(NumericTree* nt_, int cls_, cudaStream_t s_ = nullptr) : nt(nt_), cls(cls_), st(s_ ? s_ : nt_->stream) {
      if (!nt->profile) return;
      cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a, st);
   }
   ~ProfScope() {
      if (!nt->profile) return;
      cudaEventRecord(b, st);
      nt->prof_events.push_back({cls + 256 * nt->prof_level, {a, b}});
   }
};
}  // namespace



// does rank r hold (and work on) front f?  split fronts: every member of the group
static inline bool in_dest(const NumericTree* nt, int f, int r) {
   if (!nt->splitP.empty() && nt->splitP[f] > 1) return r >= nt->grp0[f] && r < nt->grp0[f] + nt->splitP[f];
   return nt->owner[f] == r;
}

// engine_indef.cu
void run_indef(NumericTree* nt, sylver_inform_c* stats);
void indef_setup(NumericTree* nt);
void indef_destroy(NumericTree* nt);
void plan_contrib_arena(NumericTree* nt);
void plan_owned_levels(NumericTree* nt, bool upload = true);
void posdef_plan_host(NumericTree* nt, bool device);
void upload_geometry(NumericTree* nt);
void load_values(NumericTree* nt, const double* aval, const double* scaling);

}  // namespace sylver_b200
