// B200 numeric factorization engine: plan construction, device arenas, the
// level-batched CUDA stream/graph scheduler and the triangular solves.
//
// Scheduling model (replaces StarPU task submission, reference
// src/NumericTree.hxx:186-406 / src/NumericTreePosdef.hxx:133-353):
//   fronts are grouped by height in the assembly tree; every kernel launch is
//   *batched* over all fronts of a level (tree parallelism inside a launch),
//   block columns of the same index advance together (node parallelism inside
//   a launch), and for positive definite problems the whole launch sequence is
//   captured once into a CUDA graph and replayed per factorization.
#include "engine.hpp"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "engine_impl.hpp"
#include "kernels_solve.cuh"

namespace sylver_b200 {

int device_count() {
   int n = 0;
   if (cudaGetDeviceCount(&n) != cudaSuccess) {
      cudaGetLastError();
      return 0;
   }
   return n;
}


static cudaStream_t g_user_stream = nullptr;
static bool g_have_user_stream = false;
void set_user_stream(void* s, bool enable) {
   g_user_stream = (cudaStream_t)s;
   g_have_user_stream = enable;
}

// ===========================================================================
// SymbolicTree
// ===========================================================================
SymbolicTree::~SymbolicTree() {
   if (!on_device) return;
   cudaFree(d_rlist); cudaFree(d_rptr); cudaFree(d_nlist); cudaFree(d_anode); cudaFree(d_nrow);
   cudaFree(d_ncol); cudaFree(d_parent); cudaFree(d_nchild); cudaFree(d_cmap); cudaFree(d_cmapoff);
   cudaFree(d_level_nodes); cudaFree(d_nptr);
   cudaFree(d_fchild); cudaFree(d_pinvoff); cudaFree(d_pinv);
}

// Host-only copies needed again at upload time
// Point-to-point operations of one rank in one NCCL group (contribution pieces).  Groups of ~400
// per rank (lap27_100 on 8 GPUs) have always worked, ~800-1050 (lap7_150) hang; smaller groups
// cost time (lap27_100 on 8 GPUs: 90 ms per step with one group per level, 102 ms with groups of
// 128, 105 ms with groups of 48 -- the transfers of a group overlap, consecutive groups do not).
constexpr int EX_MAX_OPS = 400;

struct SymbolicExtra {
   std::vector<long> rptr1;   // 1-based rptr as given
   std::vector<long> nlist;   // (src,dest) pairs as given
   std::vector<int> anode;    // owning front of each entry
   std::vector<long> nptr1;   // 1-based nptr as given
};
// rank threads of a local fabric analyse and factorize concurrently: the registry is locked
static std::mutex g_extras_mu;
static std::map<const SymbolicTree*, SymbolicExtra>& extras_map() {
   static std::map<const SymbolicTree*, SymbolicExtra> m;
   return m;
}
static SymbolicExtra& extras_of(const SymbolicTree* st) {
   std::lock_guard<std::mutex> lk(g_extras_mu);
   return extras_map()[st];      // std::map references stay valid across other insertions
}

// ===========================================================================
// Chain coarsening of the assembly tree (engine-internal; the analyse output is untouched).
//
// The reference's supernode rule (spral/src/core_analyse.f90:806-819) only merges a child
// into its parent when the column counts nest exactly, so e.g. the separator planes of a
// 7-point Laplacian arrive as chains of 25-100 thin fronts (n = one grid line, m = the whole
// separator): every link rewrites a (m-n)^2 contribution block with a rank-n update -- the
// DMMA tiles run at K = n and the tree gets hundreds of levels deep.  The engine merges a
// front into its parent when the front is the parent's LAST child (their columns are then
// adjacent in the pivot order) and the explicit zeros this adds to the child's columns (rows
// of the parent the child does not have) stay below `tol` of the merged panel.  The merged
// front factorizes the same matrix; the pivot order inside it is unchanged for the positive
// definite path.  Node numbers of the coarse tree are what the engine-level calls use;
// node_map gives reference node -> coarse node.
// ===========================================================================
namespace {
struct CoarseTree {
   int nnodes = 0;
   std::vector<int> sptr, sparent, rlist, node_map;
   std::vector<long> rptr, nptr, nlist;
};

bool coarsen_chains(int nnodes, const int* sptr, const int* sparent, const long* rptr, const int* rlist,
                    const long* nptr, const long* nlist, double tol, CoarseTree& out) {
   std::vector<int> first, last;      // groups of consecutive reference nodes [first, last]
   out.node_map.assign(nnodes, 0);
   for (int a = 0; a < nnodes;) {
      int b = a;
      long cur_n = sptr[a + 1] - sptr[a];
      long cur_k = (rptr[a + 1] - rptr[a]) - cur_n;
      while (b + 1 < nnodes && sparent[b] - 1 == b + 1) {
         const int p = b + 1;
         const long pm = rptr[p + 1] - rptr[p], pn = sptr[p + 1] - sptr[p];
         if (pm < cur_k) return false;      // malformed structure
         const double zeros = (double)cur_n * (double)(pm - cur_k);
         const double panel = (double)(cur_n + pm) * (double)(cur_n + pn);
         if (zeros > tol * panel) break;
         cur_n += pn;
         cur_k = pm - pn;
         b = p;
      }
      for (int i = a; i <= b; ++i) out.node_map[i] = (int)first.size();
      first.push_back(a);
      last.push_back(b);
      a = b + 1;
   }
   const int G = (int)first.size();
   if (G == nnodes) return false;      // nothing merged: keep the arrays as given
   out.nnodes = G;
   out.sptr.resize(G + 1);
   out.sparent.resize(G);
   out.rptr.resize(G + 1);
   out.nptr.resize(G + 1);
   out.rptr[0] = 1;
   out.nptr[0] = nptr[0];
   const long nent = nptr[nnodes] - 1;
   out.nlist.resize(2 * (size_t)nent);
   for (int g = 0; g < G; ++g) {
      const int a = first[g], b = last[g];
      out.sptr[g] = sptr[a];
      const int pb = sparent[b] - 1;
      out.sparent[g] = (pb < nnodes ? out.node_map[pb] : G) + 1;
      const int ncols = sptr[b + 1] - sptr[a];
      const int nb_b = sptr[b + 1] - sptr[b];
      const int* brows = rlist + (rptr[b] - 1) + nb_b;              // contribution rows of the top front
      const int kb = (int)(rptr[b + 1] - rptr[b]) - nb_b;
      for (int c = sptr[a]; c < sptr[b + 1]; ++c) out.rlist.push_back(c);
      out.rlist.insert(out.rlist.end(), brows, brows + kb);
      const long gm = ncols + kb;
      out.rptr[g + 1] = out.rptr[g] + gm;
      out.nptr[g + 1] = nptr[b + 1];
      for (int j = a; j <= b; ++j) {
         const long mj = rptr[j + 1] - rptr[j];
         const int nj = sptr[j + 1] - sptr[j];
         const int coff = sptr[j] - sptr[a];
         const int* jr = rlist + (rptr[j] - 1);
         for (long e = nptr[j] - 1; e < nptr[j + 1] - 1; ++e) {
            const long dest = nlist[2 * e + 1] - 1;
            const long col = dest / mj, row = dest - col * mj;
            long nrow_new;
            if (a == b) {
               nrow_new = row;
            } else if (row < nj) {
               nrow_new = coff + row;
            } else {
               const int gidx = jr[row];                            // 1-based variable index
               if (gidx < sptr[b + 1]) nrow_new = gidx - sptr[a];
               else {
                  const int* it = std::lower_bound(brows, brows + kb, gidx);
                  if (it == brows + kb || *it != gidx) return false;
                  nrow_new = ncols + (it - brows);
               }
            }
            out.nlist[2 * e] = nlist[2 * e];
            out.nlist[2 * e + 1] = (coff + col) * gm + nrow_new + 1;
         }
      }
   }
   out.sptr[G] = sptr[nnodes];
   return true;
}

// per-edge assembly maps (two-pointer merge of sorted row lists); false on a malformed tree
bool edge_maps(int nnodes, const std::vector<int>& nrow, const std::vector<int>& ncol, const std::vector<int>& parent,
               const std::vector<long>& rptr0, const std::vector<int>& rlist, std::vector<long>& cmapoff,
               std::vector<int>& cmap) {
   cmapoff.assign(nnodes + 1, 0);
   for (int i = 0; i < nnodes; ++i) cmapoff[i + 1] = cmapoff[i] + (nrow[i] - ncol[i]);
   cmap.resize(cmapoff[nnodes]);
   for (int c = 0; c < nnodes; ++c) {
      const int p = parent[c];
      const int k = nrow[c] - ncol[c];
      if (k == 0) continue;
      if (p >= nnodes) {   // root with a contribution: only legal for stand-alone dense fronts
         for (int i = 0; i < k; ++i) cmap[cmapoff[c] + i] = -1;
         continue;
      }
      const int* cr = &rlist[rptr0[c] + ncol[c]];
      const int* pr = &rlist[rptr0[p]];
      const int pm = nrow[p];
      int q = 0;
      for (int i = 0; i < k; ++i) {
         while (q < pm && pr[q] < cr[i]) ++q;
         if (q >= pm || pr[q] != cr[i]) return false;
         cmap[cmapoff[c] + i] = q;
      }
   }
   return true;
}
}  // namespace

SymbolicTree* symbolic_tree_create(int n, int nnodes, const int* sptr, const int* sparent,
                                   const long* rptr, const int* rlist, const long* nptr,
                                   const long* nlist, int* flag) {
   *flag = 0;
   SymbolicTree* st = new SymbolicTree();
   st->n = n;
   // ---- outputs defined on the REFERENCE structure: flop count, assembly maps ----
   st->ref_nnodes = nnodes;
   {
      std::vector<int> rn(nnodes), rc(nnodes), rp(nnodes);
      std::vector<long> r0(nnodes + 1);
      long flops = 0;
      for (int i = 0; i <= nnodes; ++i) r0[i] = rptr[i] - 1;
      for (int i = 0; i < nnodes; ++i) {
         rn[i] = (int)(rptr[i + 1] - rptr[i]);
         rc[i] = sptr[i + 1] - sptr[i];
         rp[i] = std::min(sparent[i] - 1, nnodes);
         if (rp[i] <= i) { *flag = SYLVER_ERROR_UNKNOWN; delete st; return nullptr; }
         const long mm = rn[i] - rc[i];
         for (long j = 1; j <= rc[i]; ++j) flops += (mm + j) * (mm + j);
         st->ref_maxfront = std::max(st->ref_maxfront, rn[i]);
      }
      st->num_flops = flops;
      std::vector<int> rl(rlist, rlist + r0[nnodes]);
      if (!edge_maps(nnodes, rn, rc, rp, r0, rl, st->ref_cmapoff, st->ref_cmap)) {
         *flag = SYLVER_ERROR_UNKNOWN; delete st; return nullptr;
      }
   }
   // ---- engine structure: chains coarsened (SYLVER_B200_AMALGAMATE=<tol>, 0 disables) ----
   CoarseTree ct;
   {
      const char* ae = getenv("SYLVER_B200_AMALGAMATE");
      const double tol = ae ? atof(ae) : 0.02;
      if (tol > 0 && nnodes > 1 && coarsen_chains(nnodes, sptr, sparent, rptr, rlist, nptr, nlist, tol, ct)) {
         st->node_map = ct.node_map;
         nnodes = ct.nnodes;
         sptr = ct.sptr.data(); sparent = ct.sparent.data(); rptr = ct.rptr.data(); rlist = ct.rlist.data();
         nptr = ct.nptr.data(); nlist = ct.nlist.data();
      } else {
         st->node_map.resize(nnodes);
         for (int i = 0; i < nnodes; ++i) st->node_map[i] = i;
      }
   }
   st->nnodes = nnodes;
   st->ref_top.assign(nnodes, 0);
   for (int i = 0; i < st->ref_nnodes; ++i) st->ref_top[st->node_map[i]] = i;
   st->nrow.resize(nnodes); st->ncol.resize(nnodes); st->parent.resize(nnodes);
   st->nchild.assign(nnodes + 1, 0); st->level.assign(nnodes + 1, 0);
   st->rptr.resize(nnodes + 1);
   for (int i = 0; i <= nnodes; ++i) st->rptr[i] = rptr[i] - 1;
   st->rlist.assign(rlist, rlist + st->rptr[nnodes]);
   for (int i = 0; i < nnodes; ++i) {
      st->nrow[i] = (int)(rptr[i + 1] - rptr[i]);
      st->ncol[i] = sptr[i + 1] - sptr[i];
      st->parent[i] = std::min(sparent[i] - 1, nnodes);
      if (st->parent[i] <= i) { *flag = SYLVER_ERROR_UNKNOWN; delete st; return nullptr; }
      st->nchild[st->parent[i]]++;
   }
   // children lists, decreasing node index (reference src/SymbolicTree.cxx:51-55)
   st->child_ptr.assign(nnodes + 2, 0);
   for (int i = 0; i <= nnodes; ++i) st->child_ptr[i + 1] = st->child_ptr[i] + st->nchild[i];
   st->child_list.resize(st->child_ptr[nnodes + 1]);
   {
      std::vector<int> fill(st->child_ptr.begin(), st->child_ptr.end() - 1);
      for (int i = nnodes - 1; i >= 0; --i) st->child_list[fill[st->parent[i]]++] = i;
   }
   if (!edge_maps(nnodes, st->nrow, st->ncol, st->parent, st->rptr, st->rlist, st->cmapoff, st->cmap)) {
      *flag = SYLVER_ERROR_UNKNOWN; delete st; return nullptr;
   }
   // fused extend-add: the two children with the largest generated elements of every front
   st->fchild.assign(2 * (size_t)nnodes, -1);
   st->pinvoff.assign(2 * (size_t)nnodes, 0);
   {
      const char* fe = getenv("SYLVER_B200_FUSE_ASM");
      const bool fuse = !(fe && fe[0] == '0');
      for (int p = 0; fuse && p < nnodes; ++p) {
         const int kp = st->nrow[p] - st->ncol[p];
         if (kp == 0) continue;
         int best[2] = {-1, -1};
         for (int ci = st->child_ptr[p]; ci < st->child_ptr[p + 1]; ++ci) {
            const int c = st->child_list[ci];
            const int k = st->nrow[c] - st->ncol[c];
            if (k < 16) continue;      // tiny blocks: the scatter kernel is as good
            if (best[0] < 0 || k > st->nrow[best[0]] - st->ncol[best[0]]) { best[1] = best[0]; best[0] = c; }
            else if (best[1] < 0 || k > st->nrow[best[1]] - st->ncol[best[1]]) best[1] = c;
         }
         for (int s = 0; s < 2; ++s) {
            const int c = best[s];
            if (c < 0) continue;
            const int k = st->nrow[c] - st->ncol[c];
            const int* cm = &st->cmap[st->cmapoff[c]];
            // rows of the child's block that land in the parent's contribution block
            int first = 0;
            while (first < k && cm[first] < st->ncol[p]) ++first;
            if (first == k) continue;
            st->fchild[2 * (size_t)p + s] = c;
            st->pinvoff[2 * (size_t)p + s] = (long)st->pinv.size();
            st->pinv.resize(st->pinv.size() + kp, -1);
            int* pv = st->pinv.data() + st->pinvoff[2 * (size_t)p + s];
            for (int i = first; i < k; ++i) pv[cm[i] - st->ncol[p]] = i;
         }
      }
   }
   // levels = height above the leaves
   for (int i = 0; i < nnodes; ++i) {
      const int p = st->parent[i];
      st->level[p] = std::max(st->level[p], st->level[i] + 1);
   }
   int nlev = 0;
   for (int i = 0; i < nnodes; ++i) nlev = std::max(nlev, st->level[i] + 1);
   st->nlevels = nlev;
   st->level_ptr.assign(nlev + 1, 0);
   for (int i = 0; i < nnodes; ++i) st->level_ptr[st->level[i] + 1]++;
   for (int l = 0; l < nlev; ++l) st->level_ptr[l + 1] += st->level_ptr[l];
   st->level_nodes.resize(nnodes);
   {
      std::vector<int> fill(st->level_ptr.begin(), st->level_ptr.end() - 1);
      for (int i = 0; i < nnodes; ++i) st->level_nodes[fill[st->level[i]]++] = i;
      for (int l = 0; l < nlev; ++l)
         std::stable_sort(st->level_nodes.begin() + st->level_ptr[l], st->level_nodes.begin() + st->level_ptr[l + 1],
                          [&](int a, int b) { return st->ncol[a] > st->ncol[b]; });
   }
   // A -> front map
   SymbolicExtra& ex = extras_of(st);
   ex.rptr1.assign(rptr, rptr + nnodes + 1);
   ex.nptr1.assign(nptr, nptr + nnodes + 1);
   st->nent = nnodes ? nptr[nnodes] - 1 : 0;
   ex.nlist.assign(nlist, nlist + 2 * st->nent);
   ex.anode.resize(st->nent);
   for (int i = 0; i < nnodes; ++i)
      for (long e = nptr[i] - 1; e < nptr[i + 1] - 1; ++e) ex.anode[e] = i;
   st->aent.resize(nnodes);
   for (int i = 0; i < nnodes; ++i) st->aent[i] = nptr[i + 1] - nptr[i];
   st->nval = 0;
   for (long e = 0; e < st->nent; ++e) st->nval = std::max(st->nval, nlist[2 * e]);
   return st;
}

static void symbolic_tree_upload(SymbolicTree* st) {
   if (st->on_device) return;
   SymbolicExtra& ex = extras_of(st);
   CU_TRY(cudaGetDevice(&st->device));
   st->d_rlist = dev_upload(st->rlist);
   st->d_rptr = dev_upload(ex.rptr1);
   st->d_nlist = dev_upload(ex.nlist);
   st->d_nptr = dev_upload(ex.nptr1);
   st->d_anode = dev_upload(ex.anode);
   st->d_nrow = dev_upload(st->nrow);
   st->d_ncol = dev_upload(st->ncol);
   st->d_parent = dev_upload(st->parent);
   st->d_nchild = dev_upload(st->nchild);
   st->d_cmap = dev_upload(st->cmap);
   st->d_cmapoff = dev_upload(st->cmapoff);
   st->d_level_nodes = dev_upload(st->level_nodes);
   st->d_fchild = dev_upload(st->fchild);
   st->d_pinvoff = dev_upload(st->pinvoff);
   st->d_pinv = dev_upload(st->pinv);
   st->on_device = true;
   ex.nlist.clear(); ex.nlist.shrink_to_fit();
   ex.anode.clear(); ex.anode.shrink_to_fit();
}

void symbolic_tree_forget(const SymbolicTree* st) {
   std::lock_guard<std::mutex> lk(g_extras_mu);
   extras_map().erase(st);
}

bool numeric_tree_posdef(const NumericTree* nt) { return nt->posdef; }

// contribution arena: allocate a level's blocks, then release its children's.  Static even
// with delayed pivots (m - n never changes).
void plan_contrib_arena(NumericTree* nt) {
   SymbolicTree* st = nt->st;
   const int N = st->nnodes;
   const int me = nt->rank;
   if ((int)nt->owner.size() != N) nt->owner.assign(N, 0);
   nt->ldc.resize(N);
   nt->coff.assign(N, 0);
   for (int f = 0; f < N; ++f) nt->ldc[f] = round_up(std::max(st->nrow[f] - st->ncol[f], 1), 4);
   SegAlloc sa;
   for (int l = 0; l < st->nlevels; ++l) {
      // blocks this rank holds from level l: its own fronts' and remote children of its fronts
      for (int i = st->level_ptr[l]; i < st->level_ptr[l + 1]; ++i) {
         const int f = st->level_nodes[i];
         const long k = st->nrow[f] - st->ncol[f];
         const int p = st->parent[f];
         const bool held = in_dest(nt, f, me) || (p < N && in_dest(nt, p, me));
         if (k > 0 && held) nt->coff[f] = sa.alloc((long)nt->ldc[f] * k);
      }
      for (int i = st->level_ptr[l]; i < st->level_ptr[l + 1]; ++i) {
         const int f = st->level_nodes[i];
         if (!in_dest(nt, f, me)) continue;
         // children were consumed by this front's assembly
         for (int ci = st->child_ptr[f]; ci < st->child_ptr[f + 1]; ++ci) {
            const int c = st->child_list[ci];
            const long k = st->nrow[c] - st->ncol[c];
            if (k > 0) sa.release(nt->coff[c], (long)nt->ldc[c] * k);
         }
         // a block sent to a remote parent is free once the level's exchange is issued
         const int p = st->parent[f];
         const long k = st->nrow[f] - st->ncol[f];
         if (k > 0 && p < N && !in_dest(nt, p, me)) sa.release(nt->coff[f], (long)nt->ldc[f] * k);
      }
   }
   nt->C_doubles = sa.peak + 4;
}

void plan_exchanges(const SymbolicTree& st, const std::vector<int>& owner, int rank,
                    std::vector<std::vector<Xfer>>& sends, std::vector<std::vector<Xfer>>& recvs) {
   const int N = st.nnodes;
   sends.assign(st.nlevels, {});
   recvs.assign(st.nlevels, {});
   for (int l = 0; l < st.nlevels; ++l)
      for (int i = st.level_ptr[l]; i < st.level_ptr[l + 1]; ++i) {
         const int f = st.level_nodes[i];
         const int p = st.parent[f];
         if (p >= N || st.nrow[f] == st.ncol[f] || owner[f] == owner[p]) continue;
         if (owner[f] == rank) sends[l].push_back(Xfer{f, owner[p]});
         else if (owner[p] == rank) recvs[l].push_back(Xfer{f, owner[f]});
      }
}

// owned fronts per level (order of st->level_nodes preserved: ncol descending)
void plan_owned_levels(NumericTree* nt, bool upload) {
   SymbolicTree* st = nt->st;
   nt->lvl_ptr.assign(st->nlevels + 1, 0);
   nt->lvl_nodes.clear();
   for (int l = 0; l < st->nlevels; ++l) {
      for (int i = st->level_ptr[l]; i < st->level_ptr[l + 1]; ++i)
         if (nt->owner[st->level_nodes[i]] == nt->rank) nt->lvl_nodes.push_back(st->level_nodes[i]);
      nt->lvl_ptr[l + 1] = (int)nt->lvl_nodes.size();
   }
   if (!upload) return;
   if (nt->d_lvl_nodes) cudaFree(nt->d_lvl_nodes);
   nt->d_lvl_nodes = dev_upload(nt->lvl_nodes);
}

// Split fronts, the batched lists without them, and the pieces of contribution blocks that
// cross GPUs.  A block of an unsplit front is one piece held by its owner; the block of a
// split front is held tile column by tile column (the DMMA tile grid of k_gemm_batched mode 1)
// by the members of its group.  Every rank that works on the parent needs every piece.  All
// ranks enumerate (level, front, piece, destination) in the same order, so the sends and
// receives of a rank pair match up.
static void plan_split(NumericTree* nt, bool device) {
   SymbolicTree* st = nt->st;
   const int N = st->nnodes, me = nt->rank, nb = nt->nb;
   nt->fac_ptr.assign(st->nlevels + 1, 0);
   nt->fac_nodes.clear();
   nt->splits.assign(st->nlevels, {});
   nt->csends.assign(st->nlevels, {});
   nt->crecvs.assign(st->nlevels, {});
   std::vector<int> split_fronts;
   size_t stage = 0;
   long pair_seq = 0;
   std::vector<int> ex_count(std::max(nt->world, 1), 0);
   int ex_max = EX_MAX_OPS;      // SYLVER_B200_EX_MAX_OPS: tests force small groups
   {
      const char* xe = getenv("SYLVER_B200_EX_MAX_OPS");
      if (xe && atoi(xe) >= 1) ex_max = atoi(xe);
   }
   for (int l = 0; l < st->nlevels; ++l) {
      ++pair_seq;      // groups never span levels
      std::fill(ex_count.begin(), ex_count.end(), 0);
      for (int i = st->level_ptr[l]; i < st->level_ptr[l + 1]; ++i) {
         const int f = st->level_nodes[i];
         if (nt->splitP[f] > 1) {
            if (in_dest(nt, f, me)) {
               SplitPlan sp;
               sp.f = f; sp.P = nt->splitP[f]; sp.q = nt->splitQ[f]; sp.g0 = nt->grp0[f];
               sp.slot = (int)split_fronts.size();
               split_fronts.push_back(f);
               stage = std::max(stage, (size_t)nt->m[f] * nb);
               nt->splits[l].push_back(sp);
            }
         } else if (nt->owner[f] == me) {
            nt->fac_nodes.push_back(f);
         }
         // ---- pieces of f's contribution block and who needs them ----
         const int p = st->parent[f];
         const int k = st->nrow[f] - st->ncol[f];
         if (p >= N || k == 0 || nt->world <= 1) continue;
         const int d0 = nt->splitP[p] > 1 ? nt->grp0[p] : nt->owner[p];
         const int d1 = d0 + (nt->splitP[p] > 1 ? nt->splitP[p] : 1);
         auto emit = [&](int src, long off, size_t count) {
            for (int d = d0; d < d1; ++d) {
               if (d == src) continue;
               // group id, computed identically on every rank (whether it takes part or not): a new
               // group starts when either end of the pair already has EX_MAX_OPS operations in it
               if (ex_count[src] >= ex_max || ex_count[d] >= ex_max) {
                  ++pair_seq;
                  std::fill(ex_count.begin(), ex_count.end(), 0);
               }
               ++ex_count[src];
               ++ex_count[d];
               const long seq = pair_seq;
               if (src == me) nt->csends[l].push_back(Piece{f, d, off, count, seq});
               else if (d == me) nt->crecvs[l].push_back(Piece{f, src, off, count, seq});
            }
         };
         const long ldc = nt->ldc[f];
         if (nt->splitP[f] > 1) {
            const int odd = st->ncol[f] & 1;      // the tile grid starts at the even column n & ~1
            const int TR = (k + odd + GT_BN - 1) / GT_BN;
            for (int tj = 0; tj < TR; ++tj) {
               const int c0 = std::max(0, tj * GT_BN - odd), c1 = std::min(k, (tj + 1) * GT_BN - odd);
               if (c1 <= c0) continue;
               emit(nt->grp0[f] + tj % nt->splitP[f], c0 * ldc + c0, (size_t)((c1 - c0) * ldc - c0));
            }
         } else {
            emit(nt->owner[f], 0, (size_t)(k * ldc));
         }
      }
      nt->fac_ptr[l + 1] = (int)nt->fac_nodes.size();
   }
   nt->stage_doubles = stage;
   nt->split_fronts = split_fronts;
   if (!device) return;      // host-only planning (sylver_b200_plan_split, CPU tests)
   // sub-communicators: every rank of the world walks the same (deterministic) list of groups
   if (nt->world > 1) {
      std::vector<std::pair<int, int>> groups;
      for (int f = 0; f < N; ++f)
         if (nt->splitP[f] > 1) {
            const std::pair<int, int> g(nt->grp0[f], nt->splitP[f]);
            if (std::find(groups.begin(), groups.end(), g) == groups.end()) groups.push_back(g);
         }
      for (auto& g : groups) {
         const int gid = comm_subgroup(g.first, g.second);
         if (gid == -2) throw CudaFailure{-52};
         for (auto& lv : nt->splits)
            for (SplitPlan& sp : lv)
               if (sp.g0 == g.first && sp.P == g.second) sp.gid = gid;
      }
   }
   if (nt->d_fac_nodes) cudaFree(nt->d_fac_nodes);
   nt->d_fac_nodes = dev_upload(nt->fac_nodes);
   if (!split_fronts.empty()) {
      nt->d_split_fronts = dev_upload(split_fronts);
      nt->d_splitP = dev_upload(nt->splitP);
      nt->d_splitQ = dev_upload(nt->splitQ);
      nt->stage_doubles = 0;      // panels are broadcast in place: no staging buffer any more
      CU_TRY(cudaMalloc(&nt->d_Wsplit, (size_t)nb * nb * sizeof(double)));
   }
   if (!nt->d_zero) {
      CU_TRY(cudaMalloc(&nt->d_zero, 4 * sizeof(int)));
      CU_TRY(cudaMemset(nt->d_zero, 0, 4 * sizeof(int)));
   }
   if (nt->world > 1) {
      std::vector<int> so(N);
      for (int f = 0; f < N; ++f) so[f] = in_dest(nt, f, me) ? me : (me + 1) % nt->world;
      nt->d_scat_owner = dev_upload(so);
   }
}

// Host-only part of the positive definite plan (no CUDA, no communicator): who works on which
// front, split fronts, arena offsets, exchange lists.  `device` adds the uploads and the
// sub-communicators.
void posdef_plan_host(NumericTree* nt, bool device) {
   SymbolicTree* st = nt->st;
   const int N = st->nnodes;
   nt->m.resize(N); nt->n.resize(N); nt->ldl.resize(N);
   nt->loff.resize(N);
   const int me = nt->rank;
   {
      // fronts above the partition cut whose rank group has several members are split over the
      // group when they are wide enough for the panel broadcasts to pay (SYLVER_B200_SPLIT=0
      // keeps every front on one GPU; SYLVER_B200_SPLIT_MIN = minimum fully-summed columns)
      std::vector<int> grpn;
      partition_tree(*st, nt->world, nt->owner, &nt->grp0, &grpn);
      nt->splitP.assign(N, 1);
      nt->splitQ.assign(N, 0);
      const char* se = getenv("SYLVER_B200_SPLIT");
      const char* sm = getenv("SYLVER_B200_SPLIT_MIN");
      const int split_min = sm ? atoi(sm) : 1024;
      if (nt->world > 1 && !(se && se[0] == '0'))
         for (int f = 0; f < N; ++f)
            if (grpn[f] >= 2 && st->ncol[f] >= split_min) {
               nt->splitP[f] = grpn[f];
               nt->splitQ[f] = me - nt->grp0[f];
            }
   }
   long loff = 0;
   for (int f = 0; f < N; ++f) {
      nt->m[f] = st->nrow[f];
      nt->n[f] = st->ncol[f];
      nt->ldl[f] = round_up(nt->m[f], 4);
      nt->loff[f] = loff;
      if (in_dest(nt, f, me)) loff += (long)nt->ldl[f] * nt->n[f];
   }
   nt->L_doubles = loff + 4;
   plan_contrib_arena(nt);
   plan_exchanges(*st, nt->owner, me, nt->sends, nt->recvs);
   plan_owned_levels(nt, device);
   plan_split(nt, device);
}

// Level work lists (tile-count prefix sums, assembly items) on top of posdef_plan_host;
// device = false keeps everything on the host (sylver_b200_plan_levels, CPU tests).
static void build_posdef_plan(NumericTree* nt, bool device = true) {
   SymbolicTree* st = nt->st;
   const int N = st->nnodes;
   const int nb = nt->nb;
   const int me = nt->rank;
   posdef_plan_host(nt, device);

   // work lists
   std::vector<int> prefix;
   std::vector<int2> asmw;
   size_t wmax = 1;
   nt->levels.resize(st->nlevels);
   for (int l = 0; l < st->nlevels; ++l) {
      LevelPlan& lp = nt->levels[l];
      lp.first = nt->fac_ptr[l];
      lp.count = nt->fac_ptr[l + 1] - lp.first;
      const int* fr = nt->fac_nodes.data() + lp.first;
      lp.max_children = 0;
      int maxn = 0;
      for (int i = 0; i < lp.count; ++i) {
         lp.max_children = std::max(lp.max_children, st->nchild[fr[i]]);
         if (st->nchild[fr[i]] > 0) lp.max_contrib = std::max(lp.max_contrib, nt->m[fr[i]] - nt->n[fr[i]]);
         maxn = std::max(maxn, nt->n[fr[i]]);
      }
      // assembly: q-th child of every parent of this level
      for (int q = 0; q < lp.max_children; ++q) {
         size_t off = asmw.size();
         for (int i = 0; i < lp.count; ++i) {
            const int f = fr[i];
            if (st->nchild[f] <= q) continue;
            const int c = st->child_list[st->child_ptr[f] + q];
            const int k = nt->m[c] - nt->n[c];
            for (int j0 = 0; j0 < k; j0 += 32) asmw.push_back(make_int2(c, j0));
         }
         lp.asm_work.emplace_back(off, (int)(asmw.size() - off));
      }
      // split fronts of this level: every member walks all children (it holds all their blocks)
      // and adds the columns it owns
      for (SplitPlan& sp : nt->splits[l]) {
         const int f = sp.f;
         for (int q = 0; q < st->nchild[f]; ++q) {
            const size_t off = asmw.size();
            const int c = st->child_list[st->child_ptr[f] + q];
            const int k = nt->m[c] - nt->n[c];
            for (int j0 = 0; j0 < k; j0 += 32) asmw.push_back(make_int2(c, j0));
            sp.asm_work.emplace_back(off, (int)(asmw.size() - off));
         }
      }
      // block-column steps
      const int nsteps = (maxn + nb - 1) / nb;
      lp.steps.resize(nsteps);
      for (int s = 0; s < nsteps; ++s) {
         LevelStep& ls = lp.steps[s];
         const int p0 = s * nb;
         int cnt = 0;
         while (cnt < lp.count && nt->n[fr[cnt]] > p0) ++cnt;   // sorted by n descending
         ls.cnt = cnt;
         ls.wld = round_up(std::min(nb, maxn - p0), 4);
         wmax = std::max(wmax, (size_t)cnt * ls.wld * ls.wld);
         ls.trsm_prefix = prefix.size();
         int acc = 0;
         for (int i = 0; i < cnt; ++i) {
            const int f = fr[i];
            const int pw = std::min(nb, nt->n[f] - p0);
            const int base = (p0 + pw) & ~1;
            prefix.push_back(acc);
            if (nt->m[f] > p0 + pw) acc += (nt->m[f] - base + GT_BM - 1) / GT_BM;
         }
         prefix.push_back(acc);
         ls.trsm_tiles = acc;
         ls.upd_prefix = prefix.size();
         acc = 0;
         for (int i = 0; i < cnt; ++i) {
            const int f = fr[i];
            const int pw = std::min(nb, nt->n[f] - p0);
            const int base = p0 + pw;
            prefix.push_back(acc);
            if (nt->n[f] > base) {
               const int TR = (nt->m[f] - base + GT_BM - 1) / GT_BM;
               const int TC = (nt->n[f] - base + GT_BN - 1) / GT_BN;
               for (int tj = 0; tj < TC; ++tj) acc += TR - tj;
            }
         }
         prefix.push_back(acc);
         ls.upd_tiles = acc;
         // look-ahead split of the same tiles: first tile column (TR tiles) / the rest
         ls.updn_prefix = prefix.size();
         acc = 0;
         for (int i = 0; i < cnt; ++i) {
            const int f = fr[i];
            const int base = p0 + std::min(nb, nt->n[f] - p0);
            prefix.push_back(acc);
            if (nt->n[f] > base) acc += (nt->m[f] - base + GT_BM - 1) / GT_BM;
         }
         prefix.push_back(acc);
         ls.updn_tiles = acc;
         ls.updr_prefix = prefix.size();
         acc = 0;
         for (int i = 0; i < cnt; ++i) {
            const int f = fr[i];
            const int base = p0 + std::min(nb, nt->n[f] - p0);
            prefix.push_back(acc);
            if (nt->n[f] > base) {
               const int TR = (nt->m[f] - base + GT_BM - 1) / GT_BM;
               const int TC = (nt->n[f] - base + GT_BN - 1) / GT_BN;
               for (int tj = 1; tj < TC; ++tj) acc += TR - tj;
            }
         }
         prefix.push_back(acc);
         ls.updr_tiles = acc;
         // paired updates: the first two tile columns (the next pair of block columns) / the rest
         ls.upd2n_prefix = prefix.size();
         acc = 0;
         for (int i = 0; i < cnt; ++i) {
            const int f = fr[i];
            const int base = p0 + std::min(nb, nt->n[f] - p0);
            prefix.push_back(acc);
            if (nt->n[f] > base) {
               const int TR = (nt->m[f] - base + GT_BM - 1) / GT_BM;
               const int TC = (nt->n[f] - base + GT_BN - 1) / GT_BN;
               for (int tj = 0; tj < std::min(TC, 2); ++tj) acc += TR - tj;
            }
         }
         prefix.push_back(acc);
         ls.upd2n_tiles = acc;
         ls.upd2r_prefix = prefix.size();
         acc = 0;
         for (int i = 0; i < cnt; ++i) {
            const int f = fr[i];
            const int base = p0 + std::min(nb, nt->n[f] - p0);
            prefix.push_back(acc);
            if (nt->n[f] > base) {
               const int TR = (nt->m[f] - base + GT_BM - 1) / GT_BM;
               const int TC = (nt->n[f] - base + GT_BN - 1) / GT_BN;
               for (int tj = 2; tj < TC; ++tj) acc += TR - tj;
            }
         }
         prefix.push_back(acc);
         ls.upd2r_tiles = acc;
      }
      // contribution tiles
      lp.contrib_prefix = prefix.size();
      int acc = 0;
      for (int i = 0; i < lp.count; ++i) {
         const int f = fr[i];
         prefix.push_back(acc);
         if (nt->m[f] > nt->n[f]) {
            const int base = nt->n[f] & ~1;
            const int TR = (nt->m[f] - base + GT_BM - 1) / GT_BM;
            for (int tj = 0; tj < TR; ++tj) acc += TR - tj;
         }
      }
      prefix.push_back(acc);
      lp.contrib_tiles = acc;
   }
   // algorithmic flops per kernel class (lower triangle only, multiply and add counted)
   for (int c = 0; c < KC_COUNT; ++c) nt->prof_flops[c] = 0;
   for (int f = 0; f < N; ++f) {
      if (!in_dest(nt, f, me)) continue;
      const double share = 1.0 / nt->splitP[f];      // split fronts: this member's part
      const double m = nt->m[f], n = nt->n[f];
      for (int p0 = 0; p0 < nt->n[f]; p0 += nb) {
         const double pw = std::min(nb, nt->n[f] - p0), p1 = p0 + pw;
         nt->prof_flops[KC_POTRF] += share * pw * pw * pw / 3.0;
         nt->prof_flops[KC_TRSM] += share * (m - p1) * pw * pw;
         // sum_{c=p1}^{n-1} (m - c) lower-triangle entries, 2*pw flops each
         const double cnt = (n - p1) * m - (n * (n - 1) - p1 * (p1 - 1)) / 2.0;
         nt->prof_flops[KC_UPDATE] += share * 2.0 * pw * cnt;
      }
      const double k = m - n;
      nt->prof_flops[KC_CONTRIB] += share * n * k * (k + 1);
   }
   // algorithmic BYTES of the bandwidth-bound classes, reported through the same
   // profile slots: extend-add = 8 B source + 16 B destination read-modify-write + 4 B index per
   // contributed entry (the entries the two fused children deliver through the DMMA epilogue are
   // not this kernel's); A scatter = 8 B value + 16 B (src, dest) pair + 8 B store per entry
   nt->prof_flops[KC_SCATTER] = 32.0 * (double)st->nent;
   for (int c = 0; c < N; ++c) {
      const int p = st->parent[c];
      if (p >= N || !in_dest(nt, p, me)) continue;
      const double k = nt->m[c] - nt->n[c];
      if (k <= 0) continue;
      const int* cm = &st->cmap[st->cmapoff[c]];
      double k0 = 0;
      while (k0 < k && cm[(int)k0] < nt->n[p]) ++k0;
      const bool fused = st->fchild[2 * (size_t)p] == c || st->fchild[2 * (size_t)p + 1] == c;
      double entries = k0 * k - k0 * (k0 - 1) / 2.0;
      if (!fused) entries += (k - k0) * (k - k0 + 1) / 2.0;
      nt->prof_flops[KC_ASSEMBLE] += 28.0 * entries / nt->splitP[p];
   }
   nt->W_doubles = wmax;
   if (!device) return;
   nt->d_prefix = dev_upload(prefix);
   nt->d_asm_work = dev_upload(asmw);
}

void upload_geometry(NumericTree* nt) {
   nt->d_m = dev_upload(nt->m); nt->d_n = dev_upload(nt->n);
   nt->d_ldl = dev_upload(nt->ldl); nt->d_ldc = dev_upload(nt->ldc);
   nt->d_loff = dev_upload(nt->loff); nt->d_coff = dev_upload(nt->coff);
   SymbolicTree* st = nt->st;
   DevTree& T = nt->T;
   T.m = nt->d_m; T.n = nt->d_n; T.ldl = nt->d_ldl; T.ldc = nt->d_ldc;
   T.loff = nt->d_loff; T.coff = nt->d_coff; T.cmapoff = st->d_cmapoff;
   T.parent = st->d_parent; T.nchild = st->d_nchild; T.cmap = st->d_cmap;
   T.L = nt->d_L; T.C = nt->d_C;
   T.fchild = st->d_fchild; T.pinvoff = st->d_pinvoff; T.pinv = st->d_pinv;
   T.splitP = nt->d_splitP; T.splitQ = nt->d_splitQ;
}

// Contribution blocks (factorization) or solve work vectors whose parent front lives on another
// GPU: one NCCL group per level, sends of this rank's fronts and receives for its parents'
// remote children.  Every rank walks the levels in the same order, so groups pair up.
// buffer of front f = base + off[f] (+ skip[f] if given), k = m - n doubles per column block.
static void issue_exchange(NumericTree* nt, int l, double* base, const std::vector<long>& off, const std::vector<int>* skip) {
   if (nt->world <= 1) return;
   const auto& sd = nt->sends[l];
   const auto& rv = nt->recvs[l];
   if (sd.empty() && rv.empty()) return;
   SymbolicTree* st = nt->st;
   auto count = [&](int f) -> size_t {
      const size_t k = (size_t)(st->nrow[f] - st->ncol[f]);
      return skip ? k : k * (size_t)nt->ldc[f];
   };
   int rc = comm_group_start();
   for (const Xfer& x : sd) rc |= comm_send(base + off[x.f] + (skip ? (*skip)[x.f] : 0), count(x.f), x.peer, nt->stream);
   for (const Xfer& x : rv) rc |= comm_recv(base + off[x.f] + (skip ? (*skip)[x.f] : 0), count(x.f), x.peer, nt->stream);
   rc |= comm_group_end();
   if (rc) throw CudaFailure{-52};
}

// Contribution pieces whose destination ranks differ from their holder (see plan_split).
//
// The pieces of a level are NOT put into one NCCL group: a level of lap7_150 on 8 GPUs has ~1000
// point-to-point operations per rank (lap27_100: ~400, which works), NCCL cuts such a group into
// several kernels at points that differ from rank to rank (each rank has a different list), and a
// kernel that waits for a message its peer only posts in a LATER kernel never returns -- the
// 8-GPU hang of round 1.  plan_split assigns every (piece, destination) pair a group id `seq`,
// the same on all ranks, such that no rank has more than EX_MAX_OPS operations in a group: a send
// and its receive always sit in the same, bounded group on both sides, and the groups follow each
// other in the same order everywhere.
static void issue_contrib_exchange(NumericTree* nt, int l) {
   if (nt->world <= 1) return;
   const auto& sd = nt->csends[l];
   const auto& rv = nt->crecvs[l];
   if (sd.empty() && rv.empty()) return;
   size_t is = 0, ir = 0;      // both lists are in increasing seq order
   while (is < sd.size() || ir < rv.size()) {
      const long chunk = std::min(is < sd.size() ? sd[is].seq : LONG_MAX, ir < rv.size() ? rv[ir].seq : LONG_MAX);
      int rc = comm_group_start();
      while (is < sd.size() || ir < rv.size()) {
         const bool take_send = is < sd.size() && (ir >= rv.size() || sd[is].seq <= rv[ir].seq);
         const Piece& x = take_send ? sd[is] : rv[ir];
         if (x.seq != chunk) break;
         if (take_send) { rc |= comm_send(nt->d_C + nt->coff[x.f] + x.off, x.count, x.peer, nt->stream); ++is; }
         else { rc |= comm_recv(nt->d_C + nt->coff[x.f] + x.off, x.count, x.peer, nt->stream); ++ir; }
      }
      rc |= comm_group_end();
      if (rc) throw CudaFailure{-52};
   }
}

// One front split over a rank group (top of the tree).  Block column j of the
// fully-summed part belongs to member j % P, tile column j of the contribution block to
// member j % P.  The owner factorizes its block column (potrf + panel solve) and broadcasts
// the panel (rows p0..m) to the group; every member then applies it to the block columns it
// owns.  Depth-1 look-ahead: the owner of the next block column updates that column first and
// runs its factorization + broadcast on the second stream beside the rest of the update.
static void issue_split_front(NumericTree* nt, const SplitPlan& sp, long& launches) {
   cudaStream_t s = nt->stream;
   cudaStream_t s2 = nt->stream2 ? nt->stream2 : nt->stream;
   const DevTree& T = nt->T;
   const int nb = nt->nb, f = sp.f, m = nt->m[f], n = nt->n[f], ldl = nt->ldl[f];
   const int P = sp.P, q = sp.q, me = nt->rank;
   const int NB = (n + nb - 1) / nb;
   const int* d_fr = nt->d_split_fronts + sp.slot;
   const TileBatch b{d_fr, nt->d_zero, 1};
   double* Lf = nt->d_L + nt->loff[f];
   auto assemble = [&](int part) {
      for (auto& w : sp.asm_work) {
         if (w.second == 0) continue;
         ProfScope ps(nt, KC_ASSEMBLE);
         k_assemble<<<w.second, 256, 0, s>>>(T, nt->d_asm_work + w.first, part);
         ++launches;
      }
   };
   auto panel = [&](int si) {
      const int p0 = si * nb, pw = std::min(nb, n - p0), rows = m - p0;
      const int own = sp.g0 + si % P;
      const int wld = round_up(pw, 4);
      if (me == own) {
         {
            ProfScope ps(nt, KC_POTRF, s2);
            if (nt->potrf_reg)
               k_potrf_inv_reg<<<1, PR_THREADS, 0, s2>>>(T, d_fr, si, nb, nt->d_Wsplit, wld, nt->d_fail);
            else
               k_potrf_inv_blk<<<1, PB_THREADS, PB_SMEM_BYTES, s2>>>(T, d_fr, si, nb, nt->d_Wsplit, wld, nt->d_fail);
            ++launches;
         }
         if (m > p0 + pw) {
            const int tiles = (m - ((p0 + pw) & ~1) + GT_BM - 1) / GT_BM;
            ProfScope ps(nt, KC_TRSM, s2);
            k_gemm_batched<<<gemm_grid(2, tiles), GT_THREADS, GT_SMEM_BYTES, s2>>>(T, b, 2, si, nb, nt->d_Wsplit, wld, 0, 1);
            ++launches;
         }
      }
      // The panel travels IN PLACE: every member keeps the front at the same leading dimension, so
      // the span from L(p0, p0) to L(m-1, p0+pw-1) is one contiguous piece of the arena on both
      // sides (it includes the rows above p0 of the later columns -- the zero upper triangle);
      // no pack / unpack copies around the broadcast.
      const size_t span = (size_t)(pw - 1) * ldl + (size_t)rows;
      if (comm_bcast(Lf + (size_t)p0 * ldl + p0, span, own, sp.gid, s2)) throw CudaFailure{-52};
   };
   // trailing update by block column si: tile columns tstart, tstart + tstep, ... of the grid that
   // starts at column (si + 1) * nb (tile column tj = block column si + 1 + tj)
   auto update = [&](int si, int tstart, int tstep, bool first_only) {
      const int base = (si + 1) * nb;
      if (n <= base) return;
      const int TR = (m - base + GT_BM - 1) / GT_BM, TC = (n - base + GT_BN - 1) / GT_BN;
      int tiles = 0;
      if (first_only) tiles = TR;
      else
         for (int tj = tstart; tj < TC; tj += tstep) tiles += TR - tj;
      if (tiles == 0) return;
      ProfScope ps(nt, KC_UPDATE);
      k_gemm_batched<<<gemm_grid(0, tiles), GT_THREADS, GT_SMEM_BYTES, s>>>(T, b, 0, si, nb, nullptr, 0, tstart, tstep);
      ++launches;
   };
   assemble(0);
   if (s2 != s) {
      CU_TRY(cudaEventRecord(nt->ev_next, s));
      CU_TRY(cudaStreamWaitEvent(s2, nt->ev_next, 0));
   }
   panel(0);
   if (s2 != s) CU_TRY(cudaEventRecord(nt->ev_panel, s2));
   for (int si = 0; si < NB; ++si) {
      if (s2 != s) CU_TRY(cudaStreamWaitEvent(s, nt->ev_panel, 0));      // panel si is here
      if (si + 1 >= NB) break;
      if (q == (si + 1) % P) {
         update(si, 0, 1, true);
         if (s2 != s) {
            CU_TRY(cudaEventRecord(nt->ev_next, s));
            CU_TRY(cudaStreamWaitEvent(s2, nt->ev_next, 0));
         }
      }
      panel(si + 1);
      if (s2 != s) CU_TRY(cudaEventRecord(nt->ev_panel, s2));
      int t0 = ((q - (si + 1)) % P + P) % P;
      if (t0 == 0) t0 = P;
      update(si, t0, P, false);
   }
   if (m > n) {
      const int base = n & ~1;
      const int TR = (m - base + GT_BM - 1) / GT_BM;
      int tiles = 0;
      for (int tj = q; tj < TR; tj += P) tiles += TR - tj;
      if (tiles > 0) {
         ProfScope ps(nt, KC_CONTRIB);
         k_gemm_batched<<<gemm_grid(1, tiles), GT_THREADS, GT_SMEM_BYTES, s>>>(T, b, 1, 0, nb, nullptr, 0, q, P);
         ++launches;
      }
   }
   assemble(1);
}

static void issue_posdef(NumericTree* nt) {
   SymbolicTree* st = nt->st;
   cudaStream_t s = nt->stream;
   const DevTree& T = nt->T;
   const int nb = nt->nb;
   long launches = 0;
   CU_TRY(cudaMemsetAsync(nt->d_L, 0, nt->L_doubles * sizeof(double), s));
   CU_TRY(cudaMemsetAsync(nt->d_fail, 0, sizeof(int), s));
   CU_TRY(cudaMemsetAsync(nt->d_fail + 1, 0x7f, sizeof(int), s));
   if (st->nent > 0) {
      const int blocks = (int)std::min<long>((st->nent + 255) / 256, 148 * 16);
      ProfScope ps(nt, KC_SCATTER);
      k_scatter_a<<<blocks, 256, 0, s>>>(T, st->nent, st->d_nlist, st->d_anode, st->d_nrow, st->d_ncol,
                                         nt->d_aval, nt->d_scaling, st->d_rlist, st->d_rptr,
                                         nt->world > 1 ? nt->d_scat_owner : nullptr, nt->rank);
      ++launches;
   }
   for (size_t l = 0; l < nt->levels.size(); ++l) {
      const LevelPlan& lp = nt->levels[l];
      nt->prof_level = (int)l;
      const int* d_fr = nt->d_fac_nodes + lp.first;
      if (lp.count == 0) {
         for (const SplitPlan& sp : nt->splits[l]) issue_split_front(nt, sp, launches);
         issue_contrib_exchange(nt, (int)l);
         continue;
      }
      auto assemble = [&](int part) {
         for (auto& w : lp.asm_work) {
            if (w.second == 0) continue;
            ProfScope ps(nt, KC_ASSEMBLE);
            k_assemble<<<w.second, 256, 0, s>>>(T, nt->d_asm_work + w.first, part);
            ++launches;
         }
      };
      assemble(0);      // children -> fully-summed columns
      auto potrf = [&](size_t si, cudaStream_t q) {
         const LevelStep& ls = lp.steps[si];
         ProfScope ps(nt, KC_POTRF, q);
         if (ls.wld <= 32)
            k_potrf_inv<32><<<ls.cnt, 128, PotrfCfg<32>::SMEM, q>>>(T, d_fr, (int)si, nb, nt->d_W, ls.wld, nt->d_fail);
         else if (ls.wld <= 64)
            k_potrf_inv<64><<<ls.cnt, 256, PotrfCfg<64>::SMEM, q>>>(T, d_fr, (int)si, nb, nt->d_W, ls.wld, nt->d_fail);
         else if (nt->potrf_old)
            k_potrf_inv<128><<<ls.cnt, 512, PotrfCfg<128>::SMEM, q>>>(T, d_fr, (int)si, nb, nt->d_W, ls.wld, nt->d_fail);
         else if (nt->potrf_reg)
            k_potrf_inv_reg<<<ls.cnt, PR_THREADS, 0, q>>>(T, d_fr, (int)si, nb, nt->d_W, ls.wld, nt->d_fail);
         else
            k_potrf_inv_blk<<<ls.cnt, PB_THREADS, PB_SMEM_BYTES, q>>>(T, d_fr, (int)si, nb, nt->d_W, ls.wld, nt->d_fail);
         ++launches;
      };
      auto trsm = [&](size_t si, cudaStream_t q) {
         const LevelStep& ls = lp.steps[si];
         if (ls.trsm_tiles == 0) return;
         TileBatch b{d_fr, nt->d_prefix + ls.trsm_prefix, ls.cnt};
         ProfScope ps(nt, KC_TRSM, q);
         k_gemm_batched<<<gemm_grid(2, ls.trsm_tiles), GT_THREADS, GT_SMEM_BYTES, q>>>(T, b, 2, (int)si, nb, nt->d_W, ls.wld, 0, 1);
         ++launches;
      };
      auto update = [&](size_t si, int sub, cudaStream_t q) {
         const LevelStep& ls = lp.steps[si];
         const int tiles = sub == 0 ? ls.upd_tiles : (sub == 1 ? ls.updn_tiles : ls.updr_tiles);
         if (tiles == 0) return;
         const size_t off = sub == 0 ? ls.upd_prefix : (sub == 1 ? ls.updn_prefix : ls.updr_prefix);
         TileBatch b{d_fr, nt->d_prefix + off, ls.cnt};
         ProfScope ps(nt, KC_UPDATE, q);
         k_gemm_batched<<<gemm_grid(0, tiles), GT_THREADS, GT_SMEM_BYTES, q>>>(T, b, 0, (int)si, nb, nullptr, 0, sub == 2 ? 1 : 0, 1);
         ++launches;
      };
      // Look-ahead (few, large fronts): as soon as the next block column has received its
      // update, its diagonal-block factorization and panel solve run on a second stream beside
      // the rest of the trailing update, taking the latency-bound potrf off the critical path.
      const bool lookahead = nt->stream2 && lp.steps.size() >= 2 && lp.count <= 16;
      // profiled runs bracket every launch with events: the SAME launches, but the chain of the next
      // pair is issued on the main stream too, so that a bracket holds the kernel's own time
      cudaStream_t la_stream = nt->profile ? s : nt->stream2;
      if (nt->pair_updates) {
         // Block columns are factorized in pairs: column si, a rank-nb update of column si+1
         // alone, column si+1 -- then ONE rank-2nb update of everything behind the pair (half as
         // many read-modify-write passes over the trailing panel and twice the work per tile).
         // With look-ahead the next pair is factorized on the second stream as soon as its two
         // tile columns have received the update, beside the rest of that update.
         const size_t S = lp.steps.size();
         auto upd = [&](const LevelStep& ls, size_t off, int tiles, int kstep, int knb, int tstart, cudaStream_t q) {
            if (tiles == 0) return;
            TileBatch b{d_fr, nt->d_prefix + off, ls.cnt};
            ProfScope ps(nt, KC_UPDATE, q);
            k_gemm_batched<<<gemm_grid(0, tiles), GT_THREADS, GT_SMEM_BYTES, q>>>(T, b, 0, kstep, knb, nullptr, 0, tstart, 1);
            ++launches;
         };
         auto chain = [&](size_t si, cudaStream_t q) {
            potrf(si, q);
            trsm(si, q);
            if (si + 1 < S) {
               upd(lp.steps[si], lp.steps[si].updn_prefix, lp.steps[si].updn_tiles, (int)si, nb, 0, q);
               potrf(si + 1, q);
               trsm(si + 1, q);
            }
         };
         chain(0, s);
         for (size_t si = 0; si + 2 < S; si += 2) {
            const LevelStep& l1 = lp.steps[si + 1];      // tile lists behind the pair (si, si+1)
            if (!lookahead) {
               upd(l1, l1.upd_prefix, l1.upd_tiles, (int)(si / 2), 2 * nb, 0, s);
               chain(si + 2, s);
            } else {
               cudaStream_t s2 = la_stream;
               upd(l1, l1.upd2n_prefix, l1.upd2n_tiles, (int)(si / 2), 2 * nb, 0, s);
               CU_TRY(cudaEventRecord(nt->ev_next, s));
               CU_TRY(cudaStreamWaitEvent(s2, nt->ev_next, 0));
               chain(si + 2, s2);
               CU_TRY(cudaEventRecord(nt->ev_panel, s2));
               upd(l1, l1.upd2r_prefix, l1.upd2r_tiles, (int)(si / 2), 2 * nb, 2, s);
               CU_TRY(cudaStreamWaitEvent(s, nt->ev_panel, 0));
            }
         }
      } else if (!lookahead) {
         for (size_t si = 0; si < lp.steps.size(); ++si) {
            potrf(si, s);
            trsm(si, s);
            update(si, 0, s);
         }
      } else {
         cudaStream_t s2 = la_stream;
         potrf(0, s);
         trsm(0, s);
         for (size_t si = 0; si < lp.steps.size(); ++si) {
            if (si + 1 < lp.steps.size()) {
               update(si, 1, s);
               CU_TRY(cudaEventRecord(nt->ev_next, s));
               CU_TRY(cudaStreamWaitEvent(s2, nt->ev_next, 0));
               potrf(si + 1, s2);
               trsm(si + 1, s2);
               CU_TRY(cudaEventRecord(nt->ev_panel, s2));
               update(si, 2, s);
               CU_TRY(cudaStreamWaitEvent(s, nt->ev_panel, 0));
            } else {
               update(si, 0, s);
            }
         }
      }
      if (lp.contrib_tiles > 0) {
         TileBatch b{d_fr, nt->d_prefix + lp.contrib_prefix, lp.count};
         ProfScope ps(nt, KC_CONTRIB);
         k_gemm_batched<<<gemm_grid(1, lp.contrib_tiles), GT_THREADS, GT_SMEM_BYTES, s>>>(T, b, 1, 0, nb, nullptr, 0, 0, 1);
         ++launches;
      }
      assemble(1);      // children -> contribution block (after this front's own Schur complement)
      for (const SplitPlan& sp : nt->splits[l]) issue_split_front(nt, sp, launches);
      issue_contrib_exchange(nt, (int)l);
   }
   if (nt->world > 1 && comm_allreduce_max_int(nt->d_fail, 1, s)) throw CudaFailure{-52};
   CU_TRY(cudaGetLastError());
   nt->launches = launches;
}

static void set_kernel_attributes() {
   static bool done = false;
   static std::mutex mu;
   std::lock_guard<std::mutex> lk(mu);
   if (done) return;
   CU_TRY(cudaFuncSetAttribute(k_gemm_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES));
   CU_TRY(cudaFuncSetAttribute(k_gemm_batched, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
   if (getenv("SYLVER_B200_VERBOSE")) {
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gemm_batched, GT_THREADS, GT_SMEM_BYTES);
      fprintf(stderr, "sylver_b200: k_gemm_batched resident CTAs per SM: %d\n", nb);
   }
   CU_TRY(cudaFuncSetAttribute(k_potrf_inv_blk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PB_SMEM_BYTES));
   CU_TRY(cudaFuncSetAttribute(k_potrf_inv<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PotrfCfg<128>::SMEM));
   CU_TRY(cudaFuncSetAttribute(k_potrf_inv<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PotrfCfg<64>::SMEM));
   CU_TRY(cudaFuncSetAttribute(k_potrf_inv<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PotrfCfg<32>::SMEM));
   done = true;
}

void load_values(NumericTree* nt, const double* aval, const double* scaling) {
   SymbolicTree* st = nt->st;
   // number of values referenced = max src index; the caller's array covers ptr[n]-1 entries,
   // which equals nent for a full-rank analysis.  Copy exactly what the map references.
   cudaPointerAttributes attr{};
   bool on_dev = false;
   if (cudaPointerGetAttributes(&attr, aval) == cudaSuccess)
      on_dev = (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
   else
      cudaGetLastError();
   auto t0 = std::chrono::steady_clock::now();
   CU_TRY(cudaMemcpyAsync(nt->d_aval, aval, nt->aval_count * sizeof(double),
                          on_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, nt->stream));
   if (scaling) {
      if (!nt->d_scaling) CU_TRY(cudaMalloc(&nt->d_scaling, std::max(st->n, 1) * sizeof(double)));
      CU_TRY(cudaMemcpyAsync(nt->d_scaling, scaling, st->n * sizeof(double), cudaMemcpyDefault, nt->stream));
   }
   CU_TRY(cudaStreamSynchronize(nt->stream));
   nt->t_h2d = on_dev ? 0.0 : std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

static void run_posdef(NumericTree* nt, sylver_inform_c* stats) {
   CU_TRY(cudaEventRecord(nt->ev0, nt->stream));
   if (nt->profile || !nt->graph) {
      for (auto& e : nt->prof_events) { cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second); }
      nt->prof_events.clear();
      issue_posdef(nt);      // NCCL exchanges are issued eagerly, level by level
   } else {
      CU_TRY(cudaGraphLaunch(nt->graph, nt->stream));
   }
   CU_TRY(cudaEventRecord(nt->ev1, nt->stream));
   int fail[2] = {0, 0};
   CU_TRY(cudaMemcpyAsync(fail, nt->d_fail, 2 * sizeof(int), cudaMemcpyDeviceToHost, nt->stream));
   CU_TRY(cudaStreamSynchronize(nt->stream));
   float ms = 0;
   CU_TRY(cudaEventElapsedTime(&ms, nt->ev0, nt->ev1));
   nt->t_device = ms * 1e-3;
   if (nt->profile) {
      for (int c = 0; c < KC_COUNT; ++c) { nt->prof_ms[c] = 0; nt->prof_launches[c] = 0; }
      nt->prof_level_ms.assign((size_t)nt->st->nlevels * KC_COUNT, 0.0);
      for (auto& e : nt->prof_events) {
         float t = 0;
         cudaEventElapsedTime(&t, e.second.first, e.second.second);
         nt->prof_ms[e.first & 255] += t;
         nt->prof_launches[e.first & 255]++;
         nt->prof_level_ms[(size_t)(e.first >> 8) * KC_COUNT + (e.first & 255)] += t;
      }
   }
   *stats = sylver_inform_c{};
   int maxfront = 0;
   for (int f = 0; f < nt->st->nnodes; ++f) maxfront = std::max(maxfront, nt->m[f]);
   stats->maxfront = maxfront;
   if (fail[0]) stats->flag = SYLVER_ERROR_NOT_POS_DEF;
}

NumericTree* numeric_tree_create(bool posdef, SymbolicTree* st, const double* aval, const double* scaling,
                                 const sylver_options_c* options, sylver_inform_c* stats) {
   NumericTree* nt = nullptr;
   auto w0 = std::chrono::steady_clock::now();
   try {
      if (device_count() == 0) throw CudaFailure{(int)cudaErrorNoDevice};
      set_kernel_attributes();
      symbolic_tree_upload(st);
      nt = new NumericTree();
      nt->st = st;
      nt->posdef = posdef;
      nt->opt = *options;
      nt->nb = 128;
      nt->rank = comm().rank;
      nt->world = comm().world;
      if (g_have_user_stream) {
         nt->stream = g_user_stream;
         nt->own_stream = false;
      } else {
         // own streams: the main stream sits between the latency-critical chain stream (stream2,
         // highest priority) and the bulk stream of the APTP contribution passes (stream3, lowest),
         // so that the CTAs of a short chain kernel are dispatched before the queued bulk tiles
         int plo = 0, phi = 0;
         CU_TRY(cudaDeviceGetStreamPriorityRange(&plo, &phi));
         CU_TRY(cudaStreamCreateWithPriority(&nt->stream, cudaStreamNonBlocking, (plo + phi) / 2));
      }
      {
         const char* pe = getenv("SYLVER_B200_PROFILE");
         nt->profile = pe && pe[0] == '1';
      }
      CU_TRY(cudaEventCreate(&nt->ev0));
      CU_TRY(cudaEventCreate(&nt->ev1));
      {
         const char* po = getenv("SYLVER_B200_POTRF_OLD");
         nt->potrf_old = po && po[0] == '1';
         const char* pr = getenv("SYLVER_B200_POTRF_REG");
         nt->potrf_reg = pr && pr[0] == '1';
         const char* pe = getenv("SYLVER_B200_PAIR");
         nt->pair_updates = !(pe && pe[0] == '0');
      }
      {
         const char* la = getenv("SYLVER_B200_LOOKAHEAD");
         if (!(la && la[0] == '0')) {
            int plo = 0, phi = 0;
            CU_TRY(cudaDeviceGetStreamPriorityRange(&plo, &phi));
            CU_TRY(cudaStreamCreateWithPriority(&nt->stream2, cudaStreamNonBlocking, phi));
            CU_TRY(cudaStreamCreateWithPriority(&nt->stream3, cudaStreamNonBlocking, plo));
            CU_TRY(cudaEventCreateWithFlags(&nt->ev_next, cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&nt->ev_panel, cudaEventDisableTiming));
         }
      }
      CU_TRY(cudaMalloc(&nt->d_fail, 4 * sizeof(int)));
      // values: the map references entries 1..max(src)
      nt->aval_count = (size_t)st->nval;
      CU_TRY(cudaMalloc(&nt->d_aval, std::max<size_t>(nt->aval_count, 1) * sizeof(double)));
      // the scaling buffer must exist before the launch sequence is captured: its address is a
      // kernel argument inside the CUDA graph (a tree is created with or without scaling and
      // keeps that property; api.cpp builds a new tree when it changes)
      if (scaling) CU_TRY(cudaMalloc(&nt->d_scaling, std::max(st->n, 1) * sizeof(double)));
      if (!posdef) {
         indef_setup(nt);
         load_values(nt, aval, scaling);
         run_indef(nt, stats);
         nt->t_wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
         return nt;
      }
      build_posdef_plan(nt);
      if (nt->world > 1) nt->d_owner = dev_upload(nt->owner);
      CU_TRY(cudaMalloc(&nt->d_L, nt->L_doubles * sizeof(double)));
      CU_TRY(cudaMalloc(&nt->d_C, nt->C_doubles * sizeof(double)));
      CU_TRY(cudaMalloc(&nt->d_W, nt->W_doubles * sizeof(double)));
      upload_geometry(nt);
      // capture the launch sequence once.  Multi-GPU runs over NCCL can capture it as well (the
      // exchanges and panel broadcasts become graph nodes; every rank replays the same sequence).
      // The in-process fabric (rank threads, host-side rendezvous) cannot be captured.
      // Opt-in (SYLVER_B200_GRAPH_MULTI=1): measured on 2 x B200 the replayed graph is no faster than
      // the eager issue (149.83 vs 149.87 ms on lap27_100, profiles/r2_two_gpu.md), and a graph that
      // holds NCCL nodes must be destroyed before its communicator.
      {
         const char* gm = getenv("SYLVER_B200_GRAPH_MULTI");
         nt->graph_multi = nt->world > 1 && comm().nccl != nullptr && gm && gm[0] == '1';
      }
      if (!nt->profile && (nt->world == 1 || nt->graph_multi)) {
         cudaGraph_t g = nullptr;
         CU_TRY(cudaStreamBeginCapture(nt->stream, cudaStreamCaptureModeThreadLocal));
         bool captured = true;
         try {
            issue_posdef(nt);
         } catch (CudaFailure&) {
            if (!nt->graph_multi) throw;
            captured = false;      // a collective that cannot be captured: eager issue instead
         }
         cudaError_t ce = cudaStreamEndCapture(nt->stream, &g);
         if (captured && ce == cudaSuccess && g) {
            CU_TRY(cudaGraphInstantiate(&nt->graph, g, 0));
         } else if (!nt->graph_multi) {
            CU_TRY(ce);
         } else {
            cudaGetLastError();
            fprintf(stderr, "sylver_b200: multi-GPU launch sequence not capturable (%s), issuing eagerly\n",
                    cudaGetErrorName(ce));
         }
         if (g) CU_TRY(cudaGraphDestroy(g));
      }
      load_values(nt, aval, scaling);
      run_posdef(nt, stats);
   } catch (CudaFailure& e) {
      *stats = sylver_inform_c{};
      stats->flag = (e.code == -98) ? SYLVER_ERROR_UNIMPLEMENTED
                                    : (e.code == -52 ? SYLVER_ERROR_CUBLAS_UNKNOWN : SYLVER_ERROR_CUDA_UNKNOWN);
      cudaGetLastError();
      if (nt) numeric_tree_destroy(nt);
      return nullptr;
   } catch (std::bad_alloc&) {
      *stats = sylver_inform_c{};
      stats->flag = SYLVER_ERROR_ALLOCATION;
      if (nt) numeric_tree_destroy(nt);
      return nullptr;
   }
   nt->t_wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
   return nt;
}

void numeric_tree_refactor(NumericTree* nt, const double* aval, const double* scaling,
                           const sylver_options_c* options, sylver_inform_c* stats) {
   auto w0 = std::chrono::steady_clock::now();
   if (options) nt->opt = *options;
   try {
      load_values(nt, aval, scaling);
      if (nt->posdef) run_posdef(nt, stats);
      else run_indef(nt, stats);
   } catch (CudaFailure&) {
      *stats = sylver_inform_c{};
      stats->flag = SYLVER_ERROR_CUDA_UNKNOWN;
   }
   nt->t_wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
}

void numeric_tree_destroy(NumericTree* nt) {
   if (!nt) return;
   if (nt->graph) cudaGraphExecDestroy(nt->graph);
   indef_destroy(nt);
   cudaFree(nt->d_m); cudaFree(nt->d_n); cudaFree(nt->d_ldl); cudaFree(nt->d_ldc);
   cudaFree(nt->d_loff); cudaFree(nt->d_coff); cudaFree(nt->d_L); cudaFree(nt->d_C);
   cudaFree(nt->d_W); cudaFree(nt->d_aval); cudaFree(nt->d_scaling); cudaFree(nt->d_fail);
   cudaFree(nt->d_prefix); cudaFree(nt->d_asm_work); cudaFree(nt->d_xw); cudaFree(nt->d_xwoff);
   cudaFree(nt->d_child_ptr); cudaFree(nt->d_child_list);
   cudaFree(nt->d_solve_bar); cudaFree(nt->d_solve_part); cudaFree(nt->d_xrhs);
   cudaFree(nt->d_splitP); cudaFree(nt->d_splitQ); cudaFree(nt->d_scat_owner); cudaFree(nt->d_fac_nodes);
   cudaFree(nt->d_split_fronts); cudaFree(nt->d_zero); cudaFree(nt->d_stage); cudaFree(nt->d_Wsplit);
   cudaFree(nt->d_lvl_nodes); cudaFree(nt->d_owner); cudaFree(nt->d_all_nodes); cudaFree(nt->d_xbuf); cudaFree(nt->d_xpack_off);
   if (nt->ev0) cudaEventDestroy(nt->ev0);
   if (nt->ev1) cudaEventDestroy(nt->ev1);
   if (nt->stream && nt->own_stream) cudaStreamDestroy(nt->stream);
   if (nt->stream2) cudaStreamDestroy(nt->stream2);
   if (nt->stream3) cudaStreamDestroy(nt->stream3);
   if (nt->ev_next) cudaEventDestroy(nt->ev_next);
   if (nt->ev_panel) cudaEventDestroy(nt->ev_panel);
   for (auto& e : nt->prof_events) { cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second); }
   delete nt;
}

void numeric_tree_timings(const NumericTree* nt, double* out4) {
   out4[0] = nt->t_device;
   out4[1] = nt->t_h2d;
   out4[2] = nt->t_wall;
   out4[3] = (double)nt->launches;
}

// Host-only: the positive definite plan of `rank` in a world of `world` ranks (no CUDA, no
// communicator).  out8: split fronts in the tree, split fronts this rank works on, factor arena
// bytes, contribution arena bytes, panel staging bytes, pieces sent, pieces received, largest
// number of point-to-point operations in one level.  pieces (6 longs each, at most cap/6):
// level, front (topmost reference node), peer, offset, count (doubles), direction (0 send, 1 recv)
// + 2 * exchange group id (the NCCL group the operation is issued in, the same on both sides).
int numeric_plan_split(SymbolicTree* st, int rank, int world, long* out8, int cap, long* pieces) {
   NumericTree nt;
   nt.st = st;
   nt.rank = rank;
   nt.world = world;
   nt.nb = 128;
   posdef_plan_host(&nt, false);
   for (int i = 0; i < 8; ++i) out8[i] = 0;
   for (int f = 0; f < st->nnodes; ++f)
      if (nt.splitP[f] > 1) {
         ++out8[0];
         if (in_dest(&nt, f, rank)) ++out8[1];
      }
   out8[2] = (long)(nt.L_doubles * sizeof(double));
   out8[3] = (long)(nt.C_doubles * sizeof(double));
   out8[4] = (long)(nt.stage_doubles * sizeof(double));
   int cnt = 0;
   for (int l = 0; l < st->nlevels; ++l) {
      out8[5] += (long)nt.csends[l].size();
      out8[6] += (long)nt.crecvs[l].size();
      out8[7] = std::max<long>(out8[7], (long)(nt.csends[l].size() + nt.crecvs[l].size()));
      for (int dir = 0; dir < 2; ++dir)
         for (const Piece& x : (dir == 0 ? nt.csends[l] : nt.crecvs[l])) {
            if (6 * cnt + 5 < cap) {
               long* o = pieces + 6 * cnt;
               o[0] = l; o[1] = st->ref_top[x.f]; o[2] = x.peer; o[3] = x.off; o[4] = (long)x.count; o[5] = dir + 2 * x.seq;
            }
            ++cnt;
         }
   }
   return cnt;
}

// Values as host memory: `val` itself, or a copy in `tmp` when it is a device pointer.
const double* values_on_host(const double* val, size_t count, std::vector<double>& tmp) {
   cudaPointerAttributes attr{};
   bool on_dev = false;
   if (cudaPointerGetAttributes(&attr, val) == cudaSuccess) on_dev = attr.type == cudaMemoryTypeDevice;
   else cudaGetLastError();
   if (!on_dev) return val;
   tmp.resize(count);
   if (cudaMemcpy(tmp.data(), val, count * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
   }
   return tmp.data();
}

// Host-only: the level plan of `rank` (batched fronts only; split fronts are planned by
// plan_split).  Per level one record of 4 longs (level, fronts, block-column steps,
// contribution tiles) followed by one record of 8 longs per step (fronts, trsm tiles, update
// tiles, first-tile-column tiles, rest, first-two-tile-columns tiles, rest, inverse stride).
// Returns the number of longs (may exceed cap).
long numeric_plan_levels(SymbolicTree* st, int rank, int world, long cap, long* out) {
   NumericTree nt;
   nt.st = st;
   nt.rank = rank;
   nt.world = world;
   nt.nb = 128;
   build_posdef_plan(&nt, false);
   long k = 0;
   auto put = [&](long v) { if (k < cap) out[k] = v; ++k; };
   for (size_t l = 0; l < nt.levels.size(); ++l) {
      const LevelPlan& lp = nt.levels[l];
      put((long)l); put(lp.count); put((long)lp.steps.size()); put(lp.contrib_tiles);
      for (const LevelStep& ls : lp.steps) {
         put(ls.cnt); put(ls.trsm_tiles); put(ls.upd_tiles); put(ls.updn_tiles); put(ls.updr_tiles);
         put(ls.upd2n_tiles); put(ls.upd2r_tiles); put(ls.wld);
      }
   }
   return k;
}

void numeric_tree_split_info(const NumericTree* nt, int* out3) {
   out3[0] = out3[1] = out3[2] = 0;
   for (size_t f = 0; f < nt->splitP.size(); ++f)
      if (nt->splitP[f] > 1) {
         ++out3[0];
         if (in_dest(nt, (int)f, nt->rank)) ++out3[1];
      }
   for (auto& l : nt->csends) out3[2] += (int)l.size();
}

long numeric_tree_bytes(const NumericTree* nt, long* factor_bytes, long* contrib_bytes) {
   if (factor_bytes) *factor_bytes = (long)(nt->L_doubles * sizeof(double));
   if (contrib_bytes) *contrib_bytes = (long)(nt->C_doubles * sizeof(double));
   return (long)((nt->L_doubles + nt->C_doubles + nt->W_doubles) * sizeof(double));
}

// out: KC_COUNT triples (ms, launches, algorithmic flops); returns KC_COUNT or 0 if not profiled
int numeric_tree_profile(const NumericTree* nt, double* out, int cap) {
   if (!nt->profile) return 0;
   for (int c = 0; c < KC_COUNT && 3 * c + 2 < cap; ++c) {
      out[3 * c] = nt->prof_ms[c];
      out[3 * c + 1] = (double)nt->prof_launches[c];
      out[3 * c + 2] = nt->prof_flops[c];
   }
   return KC_COUNT;
}

// per-level breakdown of the last profiled run: out[level * KC_COUNT + class] = ms; returns levels
int numeric_tree_profile_levels(const NumericTree* nt, double* out, int cap) {
   if (!nt->profile) return 0;
   const int n = (int)nt->prof_level_ms.size();
   for (int i = 0; i < n && i < cap; ++i) out[i] = nt->prof_level_ms[i];
   return n / KC_COUNT;
}

int numeric_tree_get_front_indef(const NumericTree* nt, int node, int* nelim, double* d, int* perm) {
   if (node < 0 || node >= nt->st->nnodes || nt->posdef) return -1;
   const int nn = nt->n[node];
   if (nelim) *nelim = nt->nelim[node];
   try {
      if (d) CU_TRY(cudaMemcpy(d, nt->T.D + nt->doff[node], 2 * (size_t)nn * sizeof(double), cudaMemcpyDeviceToHost));
      if (perm) CU_TRY(cudaMemcpy(perm, nt->T.perm + nt->permoff[node], (size_t)nn * sizeof(int), cudaMemcpyDeviceToHost));
   } catch (CudaFailure&) {
      return -51;
   }
   return 0;
}

int numeric_tree_get_front(const NumericTree* nt, int node, int* m, int* n, double* l, double* contrib) {
   if (node < 0 || node >= nt->st->nnodes) return -1;
   const int mm = nt->m[node], nn = nt->n[node];
   if (m) *m = mm;
   if (n) *n = nn;
   try {
      if (l)
         CU_TRY(cudaMemcpy2D(l, (size_t)mm * sizeof(double), nt->T.L + nt->loff[node],
                             (size_t)nt->ldl[node] * sizeof(double), (size_t)mm * sizeof(double), nn,
                             cudaMemcpyDeviceToHost));
      const int k = mm - nn;
      if (contrib && k > 0)
         CU_TRY(cudaMemcpy2D(contrib, (size_t)k * sizeof(double), nt->d_C + nt->coff[node],
                             (size_t)nt->ldc[node] * sizeof(double), (size_t)k * sizeof(double), k,
                             cudaMemcpyDeviceToHost));
   } catch (CudaFailure&) {
      return -51;
   }
   return 0;
}

// ===========================================================================
// Triangular solves (reference src/NumericTree.hxx:408-532,
// src/NumericTreePosdef.hxx:355-420): level-batched, one CTA per front.
// ===========================================================================
static void ensure_solve_workspace(NumericTree* nt) {
   if (nt->d_xw) return;
   SymbolicTree* st = nt->st;
   nt->xwoff.resize(st->nnodes + 1);
   long off = 0;
   for (int f = 0; f < st->nnodes; ++f) {
      nt->xwoff[f] = off;
      off += nt->m[f];
   }
   nt->xwoff[st->nnodes] = off;
   CU_TRY(cudaMalloc(&nt->d_xw, std::max<long>(off, 1) * sizeof(double)));
   nt->d_xwoff = dev_upload(nt->xwoff);
   if (!nt->d_child_ptr) nt->d_child_ptr = dev_upload(st->child_ptr);
   if (!nt->d_child_list) nt->d_child_list = dev_upload(st->child_list);
   if (!nt->d_solve_bar) {
      // multi-CTA solve kernels: one barrier counter per front slot of a launch, partial sums
      CU_TRY(cudaMalloc(&nt->d_solve_bar, 128 * sizeof(int)));
      CU_TRY(cudaMalloc(&nt->d_solve_part, (size_t)128 * SOLVE_BW * sizeof(double)));
   }
   if (nt->world > 1 && !nt->posdef && !nt->d_xbuf) {
      // replicated-x delta exchange (k_delta_pack): previous x + (value, changed) pairs
      nt->xbuf_cap = 3 * (size_t)std::max(st->n, 1);
      CU_TRY(cudaMalloc(&nt->d_xbuf, nt->xbuf_cap * sizeof(double)));
   }
   if (nt->world > 1 && nt->posdef && !nt->d_all_nodes) {
      // per-level packing offsets of every front's own variables (solve broadcasts)
      std::vector<int> off(st->nnodes);
      size_t cap = 1;
      for (int l = 0; l < st->nlevels; ++l) {
         int acc = 0;
         for (int i = st->level_ptr[l]; i < st->level_ptr[l + 1]; ++i) {
            off[i] = acc;
            acc += st->ncol[st->level_nodes[i]];
         }
         cap = std::max(cap, (size_t)acc);
      }
      nt->d_all_nodes = dev_upload(st->level_nodes);
      nt->d_xpack_off = dev_upload(off);
      nt->xbuf_cap = cap;
      CU_TRY(cudaMalloc(&nt->d_xbuf, cap * sizeof(double)));
   }
}

int numeric_tree_solve(const NumericTree* cnt, int job, int nrhs, double* x, int ldx) {
   NumericTree* nt = const_cast<NumericTree*>(cnt);
   SymbolicTree* st = nt->st;
   try {
      ensure_solve_workspace(nt);
      cudaPointerAttributes attr{};
      bool on_dev = false;
      if (cudaPointerGetAttributes(&attr, x) == cudaSuccess)
         on_dev = (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
      else
         cudaGetLastError();
      double* dx = x;
      if (!on_dev) {
         const size_t need = (size_t)ldx * nrhs;
         if (need > nt->xrhs_cap) {
            if (nt->d_xrhs) CU_TRY(cudaFree(nt->d_xrhs));
            nt->d_xrhs = nullptr;
            CU_TRY(cudaMalloc(&nt->d_xrhs, need * sizeof(double)));
            nt->xrhs_cap = need;
         }
         dx = nt->d_xrhs;
         CU_TRY(cudaMemcpyAsync(dx, x, need * sizeof(double), cudaMemcpyHostToDevice, nt->stream));
      }
      SolveArgs a{};
      a.T = nt->T;
      a.rlist = st->d_rlist;
      a.rptr = st->d_rptr;
      a.xwoff = nt->d_xwoff;
      a.xw = nt->d_xw;
      a.child_ptr = nt->d_child_ptr;
      a.child_list = nt->d_child_list;
      a.posdef = nt->posdef ? 1 : 0;
      if (!nt->posdef) {
         a.perm = nt->T.perm; a.permoff = nt->d_permoff;
         a.D = nt->T.D; a.doff = nt->d_doff;
         a.nelim = nt->d_nelim;
      }
      const bool do_fwd = (job == 0 || job == 1);
      const bool do_diag = !nt->posdef && (job == 0 || job == 2 || job == 4);
      const bool do_bwd = (job == 0 || job == 3 || job == 4);
      for (int r = 0; r < nrhs; ++r) {
         a.x = dx + (size_t)r * ldx;
         const bool multi = nt->world > 1;
         // multi-rank indefinite: publish what each rank changed after every level (see k_delta_pack)
         const bool delta = multi && !nt->posdef;
         const int nn = st->n, gn = (nn + 255) / 256;
         double* xprev = nt->d_xbuf;
         double* xd = delta ? nt->d_xbuf + nn : nullptr;
         auto sync_x = [&]() {
            k_delta_pack<<<gn, 256, 0, nt->stream>>>(nn, a.x, xprev, xd);
            if (comm_allreduce_sum(xd, 2 * (size_t)nn, nt->stream)) throw CudaFailure{-52};
            k_delta_unpack<<<gn, 256, 0, nt->stream>>>(nn, a.x, xprev, xd);
         };
         if (delta) CU_TRY(cudaMemcpyAsync(xprev, a.x, nn * sizeof(double), cudaMemcpyDeviceToDevice, nt->stream));
         const int* lptr = multi ? nt->lvl_ptr.data() : st->level_ptr.data();
         const int* d_nodes = multi ? nt->d_lvl_nodes : st->d_level_nodes;
         // levels with few (large) fronts: G CTAs per front (k_solve_*_multi), the whole launch
         // resident; the shared-device test fabric keeps one CTA per front (concurrent rank threads)
         const char* sge = getenv("SYLVER_B200_SOLVE_G");
         const int gcap = sge ? std::max(1, std::min(SOLVE_GMAX, atoi(sge))) : SOLVE_GMAX;
         auto group_size = [&](int count) {
            if (comm().fabric || count <= 0 || count > 32) return 1;
            return std::max(1, std::min(gcap, 128 / count));
         };
         auto launch_fwd = [&](const int* d_list, int count) {
            const int G = group_size(count);
            if (G > 1) {
               CU_TRY(cudaMemsetAsync(nt->d_solve_bar, 0, count * sizeof(int), nt->stream));
               k_solve_fwd_multi<<<count * G, SOLVE_THREADS, 0, nt->stream>>>(a, d_list, G, nt->d_solve_bar);
            } else {
               k_solve_fwd<<<count, SOLVE_THREADS, 0, nt->stream>>>(a, d_list);
            }
         };
         auto launch_bwd = [&](const int* d_list, int count) {
            const int G = group_size(count);
            if (G > 1) {
               CU_TRY(cudaMemsetAsync(nt->d_solve_bar, 0, count * sizeof(int), nt->stream));
               k_solve_bwd_multi<<<count * G, SOLVE_THREADS, 0, nt->stream>>>(a, d_list, G, nt->d_solve_bar, nt->d_solve_part);
            } else {
               k_solve_bwd<<<count, SOLVE_THREADS, 0, nt->stream>>>(a, d_list);
            }
         };
         if (do_fwd) {
            for (int l = 0; l < st->nlevels; ++l) {
               const int first = lptr[l], count = lptr[l + 1] - first;
               if (count > 0) launch_fwd(d_nodes + first, count);
               // update vectors of fronts whose parent lives on another GPU (rows >= n of xw)
               if (multi) issue_exchange(nt, l, nt->d_xw, nt->xwoff, &nt->n);
               if (delta) sync_x();
            }
            if (multi && !delta && !do_bwd) {
               // forward solve only: merge the per-rank pieces into one replicated vector
               k_zero_unowned<<<st->nnodes, 256, 0, nt->stream>>>(st->nnodes, nt->d_owner, nt->rank, st->d_rlist,
                                                                   st->d_rptr, st->d_ncol, a.x);
               if (comm_allreduce_sum(a.x, (size_t)st->n, nt->stream)) throw CudaFailure{-52};
            }
         }
         if (do_diag) {
            k_solve_diag<<<(st->nnodes + 7) / 8, 256, 0, nt->stream>>>(a, st->nnodes, multi ? nt->d_owner : nullptr, nt->rank);
            if (delta) sync_x();
         }
         if (do_bwd)
            for (int l = st->nlevels - 1; l >= 0; --l) {
               const int first = lptr[l], count = lptr[l + 1] - first;
               if (count > 0) launch_bwd(d_nodes + first, count);
               if (delta) sync_x();
               if (multi && !delta) {
                  // broadcast the entries solved at this level: pack (zeros for fronts of other
                  // ranks), all-reduce, unpack into the replicated x
                  const int af = st->level_ptr[l], ac = st->level_ptr[l + 1] - af;
                  size_t tot = 0;
                  for (int i = af; i < af + ac; ++i) tot += st->ncol[st->level_nodes[i]];
                  k_pack_level<<<ac, 256, 0, nt->stream>>>(nt->d_all_nodes + af, nt->d_xpack_off + af, nt->d_owner,
                                                           nt->rank, st->d_rlist, st->d_rptr, st->d_ncol, a.x,
                                                           nt->d_xbuf, 0, nullptr);
                  if (comm_allreduce_sum(nt->d_xbuf, tot, nt->stream)) throw CudaFailure{-52};
                  k_pack_level<<<ac, 256, 0, nt->stream>>>(nt->d_all_nodes + af, nt->d_xpack_off + af, nt->d_owner,
                                                           nt->rank, st->d_rlist, st->d_rptr, st->d_ncol, a.x,
                                                           nt->d_xbuf, 1, a.x);
               }
            }
      }
      CU_TRY(cudaGetLastError());
      if (!on_dev) {
         CU_TRY(cudaMemcpyAsync(x, dx, (size_t)ldx * nrhs * sizeof(double), cudaMemcpyDeviceToHost, nt->stream));
         CU_TRY(cudaStreamSynchronize(nt->stream));
      } else {
         CU_TRY(cudaStreamSynchronize(nt->stream));
      }
   } catch (CudaFailure&) {
      return SYLVER_ERROR_CUDA_UNKNOWN;
   }
   return SYLVER_SUCCESS;
}

}  // namespace sylver_b200
