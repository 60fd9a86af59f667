// Indefinite (APTP) numeric factorization: host-side level scheduler.
//
// Replaces NumericTree::factor_mf_indef + the StarPU task graph of
// FactorIndefAPP (reference src/NumericTree.hxx:186-406, src/factor_indef.hxx:1484-1646,
// src/factor_failed.hxx:28-157, src/assemble.hxx:152-302).
//
// Delayed pivots change the size of the parent front at run time
// (reference NumericFront.hxx:79-95), so unlike the positive definite path the launch
// sequence is not pre-captured: levels of the assembly tree are issued one after the
// other; after each level the host reads back one int per front (its eliminated column
// count), sizes the parents (m, n, ldl, arena offsets) and uploads one packed buffer with
// the next level's geometry and work lists.  Everything inside a level -- block sizes,
// pass counts, failed-column swaps, update extents -- is decided on the device from the
// per-front FrontState; the host never looks at numerical values.
#include "engine_impl.hpp"
#include "kernels_indef.cuh"

namespace sylver_b200 {

struct GeoUpd {
   int f, m, n, ldl;
   long loff, woff, doff, permoff;
};

static __global__ void k_set_geometry(const GeoUpd* __restrict__ u, int cnt, int* m, int* n, int* ldl, long* loff,
                                      long* woff, long* doff, long* permoff) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= cnt) return;
   const GeoUpd g = u[i];
   m[g.f] = g.m; n[g.f] = g.n; ldl[g.f] = g.ldl;
   loff[g.f] = g.loff; woff[g.f] = g.woff; doff[g.f] = g.doff; permoff[g.f] = g.permoff;
}

// Publishes the eliminated-column counts of ALL fronts of a level (after the all-reduce of a
// multi-rank run they include the fronts of other ranks): the per-front array the solves
// read, and state[f].nelim, which k_assemble_delays reads for a remote child's ghost.
static __global__ void k_publish_nelim(const int* __restrict__ fronts, const int* __restrict__ lvl_nelim, int cnt,
                                       int* __restrict__ nelim_all, FrontState* __restrict__ state) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= cnt) return;
   const int f = fronts[i];
   nelim_all[f] = lvl_nelim[i];
   state[f].nelim = lvl_nelim[i];
}

// ---- factor arena: bump allocation in chunks, offsets relative to chunk 0 ----
static long arena_alloc(NumericTree* nt, size_t doubles) {
   doubles = (doubles + 15) & ~(size_t)15;
   for (;;) {
      NumericTree::Chunk& c = nt->chunks.back();
      if (c.used + doubles <= c.cap) {
         const long off = (long)((c.ptr + c.used) - nt->chunks[0].ptr);
         c.used += doubles;
         return off;
      }
      // grow: a new chunk (only reached when delays outgrow the analyse-time estimate)
      NumericTree::Chunk nc{nullptr, std::max<size_t>(doubles, std::max<size_t>(nt->chunks[0].cap / 4, (size_t)1 << 24)), 0};
      CU_TRY(cudaMalloc(&nc.ptr, nc.cap * sizeof(double)));
      nt->chunks.push_back(nc);
   }
}

template <typename T>
static void ensure_cap(T*& dptr, size_t& cap, size_t need) {
   if (need <= cap) return;
   if (dptr) CU_TRY(cudaFree(dptr));
   cap = need + need / 4 + 64;
   CU_TRY(cudaMalloc(&dptr, cap * sizeof(T)));
}

void indef_setup(NumericTree* nt) {
   // kernels.cuh kernels are per translation unit (static __global__): this TU's copy needs its
   // own opt-in to > 48 KB of dynamic shared memory
   CU_TRY(cudaFuncSetAttribute(k_gemm_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
   CU_TRY(cudaFuncSetAttribute(k_gemm_batched, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
   CU_TRY(cudaFuncSetAttribute(k_gemm_diag3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES));
   CU_TRY(cudaFuncSetAttribute(k_gemm_diag3, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
   SymbolicTree* st = nt->st;
   const int N = st->nnodes;
   nt->m.assign(N, 0); nt->n.assign(N, 0); nt->ldl.assign(N, 0); nt->loff.assign(N, 0);
   nt->woff.assign(N, 0); nt->doff.assign(N, 0); nt->permoff.assign(N, 0);
   nt->nelim.assign(N, 0);
   partition_tree(*st, nt->world, nt->owner);
   plan_contrib_arena(nt);
   plan_exchanges(*st, nt->owner, nt->rank, nt->sends, nt->recvs);
   plan_owned_levels(nt);
   if (nt->world > 1) nt->d_owner = dev_upload(nt->owner);
   CU_TRY(cudaMalloc(&nt->d_C, nt->C_doubles * sizeof(double)));
   size_t est = 0;
   for (int f = 0; f < N; ++f) {
      if (nt->owner[f] != nt->rank) continue;
      const size_t ldl = round_up(st->nrow[f], 4);
      est += ((ldl * st->ncol[f] + 2 * (size_t)st->ncol[f] + (st->ncol[f] + 1) / 2 + 15) & ~(size_t)15);
   }
   est += est / 50 + ((size_t)1 << 16);
   NumericTree::Chunk c0{nullptr, est, 0};
   CU_TRY(cudaMalloc(&c0.ptr, c0.cap * sizeof(double)));
   nt->chunks.push_back(c0);
   nt->L_doubles = est;
   CU_TRY(cudaMalloc(&nt->d_m, N * sizeof(int)));
   CU_TRY(cudaMalloc(&nt->d_n, N * sizeof(int)));
   CU_TRY(cudaMalloc(&nt->d_ldl, N * sizeof(int)));
   CU_TRY(cudaMalloc(&nt->d_loff, N * sizeof(long)));
   CU_TRY(cudaMalloc(&nt->d_woff, N * sizeof(long)));
   CU_TRY(cudaMalloc(&nt->d_doff, N * sizeof(long)));
   CU_TRY(cudaMalloc(&nt->d_permoff, N * sizeof(long)));
   CU_TRY(cudaMalloc(&nt->d_state, N * sizeof(FrontState)));
   CU_TRY(cudaMalloc(&nt->d_nelim, N * sizeof(int)));
   CU_TRY(cudaMalloc(&nt->d_stats, 8 * sizeof(int)));
   CU_TRY(cudaEventCreateWithFlags(&nt->ev_lvl, cudaEventDisableTiming));
   for (int i = 0; i < 2; ++i) CU_TRY(cudaEventCreateWithFlags(&nt->ev_cpass[i], cudaEventDisableTiming));
   CU_TRY(cudaEventCreateWithFlags(&nt->ev_bulk, cudaEventDisableTiming));
   nt->d_ldc = dev_upload(nt->ldc);
   nt->d_coff = dev_upload(nt->coff);
   nt->d_ncol0 = dev_upload(st->ncol);
   DevTree& T = nt->T;
   T.m = nt->d_m; T.n = nt->d_n; T.ldl = nt->d_ldl; T.ldc = nt->d_ldc;
   T.loff = nt->d_loff; T.coff = nt->d_coff; T.cmapoff = st->d_cmapoff;
   T.parent = st->d_parent; T.nchild = st->d_nchild; T.cmap = st->d_cmap;
   T.L = nt->chunks[0].ptr; T.C = nt->d_C;
   T.ncol0 = nt->d_ncol0; T.woff = nt->d_woff; T.doff = nt->d_doff; T.permoff = nt->d_permoff;
   T.W = nullptr; T.D = nt->chunks[0].ptr; T.perm = reinterpret_cast<int*>(nt->chunks[0].ptr);
   T.state = nt->d_state;
   T.fchild = st->d_fchild; T.pinvoff = st->d_pinvoff; T.pinv = st->d_pinv;
   for (int c = 0; c < KC_COUNT; ++c) nt->prof_flops[c] = 0;
}

void indef_destroy(NumericTree* nt) {
   for (size_t i = 0; i < nt->chunks.size(); ++i) cudaFree(nt->chunks[i].ptr);
   nt->chunks.clear();
   cudaFree(nt->d_ncol0); cudaFree(nt->d_woff); cudaFree(nt->d_doff); cudaFree(nt->d_permoff);
   cudaFree(nt->d_state); cudaFree(nt->d_nelim); cudaFree(nt->d_stats); cudaFree(nt->d_diag);
   cudaFree(nt->d_lvl); cudaFree(nt->d_lvl_out);
   if (nt->h_lvl) cudaFreeHost(nt->h_lvl);
   if (nt->h_lvl_out) cudaFreeHost(nt->h_lvl_out);
   if (nt->ev_lvl) cudaEventDestroy(nt->ev_lvl);
   for (int i = 0; i < 2; ++i) { if (nt->ev_cpass[i]) cudaEventDestroy(nt->ev_cpass[i]); nt->ev_cpass[i] = nullptr; }
   if (nt->ev_bulk) cudaEventDestroy(nt->ev_bulk);
   nt->ev_bulk = nullptr;
   nt->ev_lvl = nullptr; nt->lvl_busy = false;
   nt->d_ncol0 = nullptr; nt->d_woff = nt->d_doff = nt->d_permoff = nullptr;
   nt->d_state = nullptr; nt->d_nelim = nt->d_stats = nullptr; nt->d_diag = nullptr;
   nt->d_lvl = nullptr; nt->h_lvl = nullptr; nt->d_lvl_out = nullptr; nt->h_lvl_out = nullptr;
}

namespace {
// packs heterogeneous arrays into one upload buffer, 16 B aligned sections
struct Packer {
   std::vector<char> buf;
   template <typename T>
   size_t put(const std::vector<T>& v) {
      const size_t off = (buf.size() + 15) & ~(size_t)15;
      buf.resize(off + v.size() * sizeof(T));
      if (!v.empty()) memcpy(buf.data() + off, v.data(), v.size() * sizeof(T));
      return off;
   }
};
}  // namespace

// (re)allocates the per-level upload buffer pair and copies `pk` to the device
static void upload_level_buffer(NumericTree* nt, const Packer& pk) {
   cudaStream_t s = nt->stream;
   if (pk.buf.size() > nt->lvl_cap) {
      CU_TRY(cudaStreamSynchronize(s));
      if (nt->h_lvl) CU_TRY(cudaFreeHost(nt->h_lvl));
      char* dl = static_cast<char*>(nt->d_lvl);
      size_t cap = nt->lvl_cap;
      ensure_cap(dl, cap, pk.buf.size());
      nt->d_lvl = dl;
      nt->lvl_cap = cap;
      CU_TRY(cudaMallocHost(&nt->h_lvl, cap));
   }
   if (pk.buf.empty()) return;
   // the previous upload may still be reading the pinned mirror (and kernels the device copy)
   if (nt->lvl_busy) CU_TRY(cudaEventSynchronize(nt->ev_lvl));
   memcpy(nt->h_lvl, pk.buf.data(), pk.buf.size());
   CU_TRY(cudaMemcpyAsync(nt->d_lvl, nt->h_lvl, pk.buf.size(), cudaMemcpyHostToDevice, s));
   CU_TRY(cudaEventRecord(nt->ev_lvl, s));
   nt->lvl_busy = true;
}

// Hand-over of one level's results to the ranks that own the parents (multi-rank runs):
// contribution blocks as in the positive definite path, plus -- the part delayed pivots add,
// SURVEY.md 8e "if indefinite with delays" -- the failed columns [nelim, n) of the child's L
// panel and their slice of the pivot permutation.  The receiver keeps them in a "ghost" of
// the child (only those columns are backed by memory) so that k_assemble_delays reads a
// remote child exactly like a local one.  Both sides know every nelim (all-reduced per
// level), so message sizes need no handshake.
static void exchange_level_indef(NumericTree* nt, int l) {
   SymbolicTree* st = nt->st;
   cudaStream_t s = nt->stream;
   const auto& sd = nt->sends[l];
   const auto& rv = nt->recvs[l];
   if (sd.empty() && rv.empty()) return;
   std::vector<GeoUpd> ghosts;
   for (const Xfer& x : rv) {
      const int c = x.f;
      const int nd = nt->n[c] - nt->nelim[c];
      if (nd <= 0) continue;
      const size_t cols = (size_t)nd * nt->ldl[c];
      const long off = arena_alloc(nt, cols + (nd + 1) / 2);
      nt->loff[c] = off - (long)nt->nelim[c] * nt->ldl[c];
      nt->permoff[c] = 2 * (off + (long)cols) - nt->nelim[c];
      ghosts.push_back(GeoUpd{c, nt->m[c], nt->n[c], nt->ldl[c], nt->loff[c], 0, 0, nt->permoff[c]});
   }
   if (!ghosts.empty()) {
      Packer pk;
      pk.put(ghosts);
      upload_level_buffer(nt, pk);
      k_set_geometry<<<((int)ghosts.size() + 255) / 256, 256, 0, s>>>(static_cast<const GeoUpd*>(nt->d_lvl), (int)ghosts.size(),
                                                                     nt->d_m, nt->d_n, nt->d_ldl, nt->d_loff, nt->d_woff,
                                                                     nt->d_doff, nt->d_permoff);
   }
   double* Lb = nt->chunks[0].ptr;
   int* Pb = reinterpret_cast<int*>(nt->chunks[0].ptr);
   int rc = comm_group_start();
   for (const Xfer& x : sd) {
      const int f = x.f;
      const size_t k = (size_t)(st->nrow[f] - st->ncol[f]);
      rc |= comm_send(nt->d_C + nt->coff[f], k * nt->ldc[f], x.peer, s);
      const int ne = nt->nelim[f], nd = nt->n[f] - ne;
      if (nd > 0) {
         rc |= comm_send(Lb + nt->loff[f] + (long)ne * nt->ldl[f], (size_t)nd * nt->ldl[f], x.peer, s);
         rc |= comm_send_int(Pb + nt->permoff[f] + ne, (size_t)nd, x.peer, s);
      }
   }
   for (const Xfer& x : rv) {
      const int c = x.f;
      const size_t k = (size_t)(st->nrow[c] - st->ncol[c]);
      rc |= comm_recv(nt->d_C + nt->coff[c], k * nt->ldc[c], x.peer, s);
      const int ne = nt->nelim[c], nd = nt->n[c] - ne;
      if (nd > 0) {
         rc |= comm_recv(Lb + nt->loff[c] + (long)ne * nt->ldl[c], (size_t)nd * nt->ldl[c], x.peer, s);
         rc |= comm_recv_int(Pb + nt->permoff[c] + ne, (size_t)nd, x.peer, s);
      }
   }
   rc |= comm_group_end();
   if (rc) throw CudaFailure{-52};
}

void run_indef(NumericTree* nt, sylver_inform_c* stats) {
   SymbolicTree* st = nt->st;
   const int N = st->nnodes;
   cudaStream_t s = nt->stream;
   const double u = nt->opt.u, small = nt->opt.small;
   const bool tpp_everywhere = (nt->opt.failed_pivot_method != 2) || (nt->opt.pivot_method == 3);
   // Only pivot_method = 2 (APP block, the default) runs the a-posteriori pass; with 1 (APP
   // aggressive) and 3 (TPP) the reference's tree code goes straight to threshold partial pivoting
   // on the whole front (src/factor_indef.hxx:95-139, src/factor_failed.hxx:55-75) -- its serial,
   // "not for performance" path, and the same one-CTA kernel here.
   const bool app_block = nt->opt.pivot_method != 1 && nt->opt.pivot_method != 3;
   const bool multi = nt->world > 1;
   const int me = nt->rank;
   int OB = OB_DEFAULT;
   {
      const char* oe = getenv("SYLVER_B200_OB");
      if (oe && atoi(oe) >= IB) OB = atoi(oe) / IB * IB;
   }
   // SYLVER_B200_APTP_LEGACY=1: round-1 launch sequence (apply / finish / swap as separate launches,
   // one contribution update after the last panel) for A/B runs
   // Rank threads of the in-process fabric share ONE device: kernels whose CTAs wait for each other
   // (k_block_column32, k_tpp_multi) assume that the launch they belong to gets the GPU's CTA slots
   // in order, which several concurrently launching ranks break -- the launch sequence without
   // device-side rendezvous is used there.
   const bool shared_device = comm().fabric != nullptr;
   const bool legacy = shared_device || (getenv("SYLVER_B200_APTP_LEGACY") && getenv("SYLVER_B200_APTP_LEGACY")[0] == '1');
   // SYLVER_B200_FUSE_DIAG=0: the 32 x 32 diagonal factorization as a launch of its own (A/B runs)
   const bool fuse_diag = !(getenv("SYLVER_B200_FUSE_DIAG") && getenv("SYLVER_B200_FUSE_DIAG")[0] == '0');
   long launches = 0;
   for (auto& c : nt->chunks) c.used = 0;
   if (nt->d_xw) { cudaFree(nt->d_xw); cudaFree(nt->d_xwoff); nt->d_xw = nullptr; nt->d_xwoff = nullptr; }
   if (nt->profile) {
      for (auto& e : nt->prof_events) { cudaEventDestroy(e.second.first); cudaEventDestroy(e.second.second); }
      nt->prof_events.clear();
   }
   CU_TRY(cudaEventRecord(nt->ev0, s));
   CU_TRY(cudaMemsetAsync(nt->d_stats, 0, 8 * sizeof(int), s));
   int maxfront = 0;
   std::vector<int> order, slot;
   for (int l = 0; l < st->nlevels; ++l) {
      const int gfirst = st->level_ptr[l], gcnt = st->level_ptr[l + 1] - gfirst;
      if (gcnt == 0) continue;
      nt->prof_level = l;
      // ---- geometry of the level (children are complete: their nelim is known on every rank).
      // Sizes are computed for all fronts of the level, memory only for the fronts of this rank.
      order.clear();
      std::vector<GeoUpd> geo;
      std::vector<int2> delay_work;
      size_t wtotal = 0;
      std::vector<std::pair<size_t, size_t>> runs;      // arena ranges to clear (bytes offsets from chunk 0)
      int max_children = 0, maxn = 0, maxk = 0;
      long max_ent = 0;
      for (int i = 0; i < gcnt; ++i) {
         const int f = st->level_nodes[gfirst + i];
         const bool mine = nt->owner[f] == me;
         int nd = 0;
         for (int ci = st->child_ptr[f]; ci < st->child_ptr[f + 1]; ++ci) {
            const int c = st->child_list[ci];
            const int d = nt->n[c] - nt->nelim[c];
            if (d > 0 && mine) delay_work.push_back(make_int2(c, st->ncol[f] + nd));
            nd += d;
         }
         const int m = st->nrow[f] + nd, n = st->ncol[f] + nd;
         const int ldl = round_up(m, 4);
         nt->m[f] = m; nt->n[f] = n; nt->ldl[f] = ldl;
         maxfront = std::max(maxfront, m);
         if (!mine) {
            nt->loff[f] = nt->woff[f] = nt->doff[f] = nt->permoff[f] = 0;
            geo.push_back(GeoUpd{f, m, n, ldl, 0, 0, 0, 0});
            continue;
         }
         order.push_back(f);
         max_children = std::max(max_children, st->nchild[f]);
         if (st->nchild[f] > 0) maxk = std::max(maxk, st->nrow[f] - st->ncol[f]);
         max_ent = std::max(max_ent, st->aent[f]);
         maxn = std::max(maxn, n);
         const size_t panel = (size_t)ldl * n;
         const size_t block = panel + 2 * (size_t)n + (n + 1) / 2;
         const long off = arena_alloc(nt, block);
         nt->loff[f] = off;
         nt->doff[f] = off + (long)panel;
         nt->permoff[f] = 2 * (off + (long)panel + 2 * (long)n);
         nt->woff[f] = (long)wtotal;
         wtotal += (panel + 15) & ~(size_t)15;
         const size_t bytes = ((block + 15) & ~(size_t)15) * sizeof(double);
         if (!runs.empty() && runs.back().first + runs.back().second == (size_t)off * sizeof(double))
            runs.back().second += bytes;
         else
            runs.emplace_back((size_t)off * sizeof(double), bytes);
      }
      const int cnt = (int)order.size();
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nt->n[a] > nt->n[b]; });
      slot.resize(cnt);
      {
         // position of every owned front in the level's global list (where its nelim is published)
         std::vector<std::pair<int, int>> pos(gcnt);
         for (int i = 0; i < gcnt; ++i) pos[i] = {st->level_nodes[gfirst + i], i};
         std::sort(pos.begin(), pos.end());
         for (int i = 0; i < cnt; ++i)
            slot[i] = std::lower_bound(pos.begin(), pos.end(), std::make_pair(order[i], 0))->second;
      }
      for (int i = 0; i < cnt; ++i) {
         const int f = order[i];
         geo.push_back(GeoUpd{f, nt->m[f], nt->n[f], nt->ldl[f], nt->loff[f], nt->woff[f], nt->doff[f], nt->permoff[f]});
      }
      // work lists (prefix sums are nested: step s uses the first cnt_s fronts of `order`)
      std::vector<int> row_prefix(cnt + 1), upd_prefix(cnt + 1), con_prefix(cnt + 1), inn_prefix(cnt + 1);
      {
         int a3 = 0;
         for (int i = 0; i < cnt; ++i) {
            // block column -> rest of its outer panel: at most OB / 128 + 1 tile columns
            inn_prefix[i] = a3;
            a3 += (OB / GT_BN + 1) * ((nt->m[order[i]] + GT_BM - 1) / GT_BM + 1);
         }
         inn_prefix[cnt] = a3;
      }
      {
         int a0 = 0, a1 = 0, a2 = 0;
         for (int i = 0; i < cnt; ++i) {
            const int f = order[i];
            const int m = nt->m[f], n = nt->n[f];
            row_prefix[i] = a0; upd_prefix[i] = a1; con_prefix[i] = a2;
            a0 += 1 + (m + AP_THREADS - 1) / AP_THREADS;
            const int TR = (m + GT_BM - 1) / GT_BM, TC = (n + GT_BN - 1) / GT_BN;
            for (int tj = 0; tj < TC; ++tj) a1 += TR - tj;
            if (m > n) {
               const int base = n & ~1;
               const int TRc = (m - base + GT_BM - 1) / GT_BM;
               a2 += TRc * (TRc + 1) / 2;
            }
         }
         row_prefix[cnt] = a0; upd_prefix[cnt] = a1; con_prefix[cnt] = a2;
      }
      std::vector<int2> asmw;
      std::vector<std::pair<size_t, int>> asm_ranges;
      for (int q = 0; q < max_children; ++q) {
         const size_t off = asmw.size();
         for (int i = 0; i < cnt; ++i) {
            const int f = order[i];
            if (st->nchild[f] <= q) continue;
            const int c = st->child_list[st->child_ptr[f] + q];
            const int k = st->nrow[c] - st->ncol[c];
            for (int j0 = 0; j0 < k; j0 += 32) asmw.push_back(make_int2(c, j0));
         }
         asm_ranges.emplace_back(off, (int)(asmw.size() - off));
      }
      Packer pk;
      const size_t o_geo = pk.put(geo), o_order = pk.put(order), o_row = pk.put(row_prefix),
                   o_upd = pk.put(upd_prefix), o_con = pk.put(con_prefix), o_asm = pk.put(asmw),
                   o_del = pk.put(delay_work), o_slot = pk.put(slot), o_inn = pk.put(inn_prefix);
      // ---- device buffers for the level ----
      {
         if ((size_t)gcnt > nt->lvl_out_cap) {
            CU_TRY(cudaStreamSynchronize(s));
            if (nt->h_lvl_out) CU_TRY(cudaFreeHost(nt->h_lvl_out));
            ensure_cap(nt->d_lvl_out, nt->lvl_out_cap, (size_t)gcnt);
            CU_TRY(cudaMallocHost(&nt->h_lvl_out, nt->lvl_out_cap * sizeof(int)));
         }
         if ((size_t)cnt > nt->diag_cap) {
            CU_TRY(cudaStreamSynchronize(s));
            DiagScratch* dd = static_cast<DiagScratch*>(nt->d_diag);
            ensure_cap(dd, nt->diag_cap, (size_t)cnt);
            nt->d_diag = dd;
         }
         if (wtotal > nt->Wscratch_cap) {
            CU_TRY(cudaStreamSynchronize(s));
            ensure_cap(nt->d_W, nt->Wscratch_cap, wtotal);
            nt->W_doubles = nt->Wscratch_cap;
         }
      }
      nt->T.W = nt->d_W;
      const DevTree& T = nt->T;
      upload_level_buffer(nt, pk);
      char* dl = static_cast<char*>(nt->d_lvl);
      const GeoUpd* d_geo = reinterpret_cast<const GeoUpd*>(dl + o_geo);
      const int* d_fr = reinterpret_cast<const int*>(dl + o_order);
      const int* d_row = reinterpret_cast<const int*>(dl + o_row);
      const int* d_upd = reinterpret_cast<const int*>(dl + o_upd);
      const int* d_con = reinterpret_cast<const int*>(dl + o_con);
      const int2* d_asm = reinterpret_cast<const int2*>(dl + o_asm);
      const int2* d_del = reinterpret_cast<const int2*>(dl + o_del);
      const int* d_slot = reinterpret_cast<const int*>(dl + o_slot);
      const int* d_inn = reinterpret_cast<const int*>(dl + o_inn);
      DiagScratch* d_diag = static_cast<DiagScratch*>(nt->d_diag);
      k_set_geometry<<<((int)geo.size() + 255) / 256, 256, 0, s>>>(d_geo, (int)geo.size(), nt->d_m, nt->d_n, nt->d_ldl,
                                                                   nt->d_loff, nt->d_woff, nt->d_doff, nt->d_permoff);
      if (multi) CU_TRY(cudaMemsetAsync(nt->d_lvl_out, 0, gcnt * sizeof(int), s));
      for (auto& r : runs)
         CU_TRY(cudaMemsetAsync(reinterpret_cast<char*>(nt->chunks[0].ptr) + r.first, 0, r.second, s));
      launches += 1 + (long)runs.size();
      if (cnt > 0) {
         // ---- assembly: A entries, children's generated elements, delayed columns ----
         {
            ProfScope ps(nt, KC_SCATTER);
            k_init_front<<<cnt, 256, 0, s>>>(T, d_fr, st->d_rlist, st->d_rptr);
            const int gy = (int)std::min<long>(std::max<long>((max_ent + 2047) / 2048, 1), 592);
            k_scatter_a_fronts<<<dim3(cnt, gy), 256, 0, s>>>(T, d_fr, st->d_nptr, st->d_nlist, st->d_nrow, nt->d_aval,
                                                             nt->d_scaling, st->d_rlist, st->d_rptr);
            launches += 2;
         }
         auto assemble = [&](int part) {
            ProfScope ps(nt, KC_ASSEMBLE);
            for (auto& w : asm_ranges) {
               if (w.second == 0) continue;
               k_assemble_indef<<<w.second, 256, 0, s>>>(T, d_asm + w.first, part);
               ++launches;
            }
         };
         assemble(0);
         if (!delay_work.empty()) {
            k_assemble_delays<<<(int)delay_work.size(), 256, 0, s>>>(T, d_del);
            ++launches;
         }
         // ---- APTP, two-level blocking: outer panels of OB candidates, IB-wide block columns
         // inside.  After o panels a front has max(0, n - o*OB) active candidates left and inside
         // a panel every block column consumes IB of them (passed or failed), so the launch
         // sequence is known to the host although the outcome of the pivoting is not.
         const int nouter = app_block ? (maxn + OB - 1) / OB : 0;
         // Every completed outer panel updates the contribution block with its own pivots right
         // away (mode 6), on the second stream: that pass only reads rows >= n of the panel's L and
         // W columns, which nothing touches afterwards, so it runs beside the pivoting chain of the
         // next panel instead of as one big update after the last one.  The panel's pivot range
         // is published per parity slot (k_outer_end); a slot is reused two panels later.
         const char* s2e = getenv("SYLVER_B200_APTP_S2");
         cudaStream_t s2 = (nt->stream3 && !legacy && !(s2e && s2e[0] == '0')) ? nt->stream3 : s;
         // SYLVER_B200_APTP_BULK1=1: bulk launches beside the pivoting chain ask for enough shared
         // memory that only ONE of their CTAs fits per SM, leaving room for the chain's kernels
         const char* b1e = getenv("SYLVER_B200_APTP_BULK1");
         const size_t bulk_smem = (s2 != s && b1e && b1e[0] == '1') ? (size_t)120 * 1024 : GT_SMEM_BYTES;
         bool cpass[2] = {false, false};
         bool bulk = false;
         int cnt_o = cnt;
         for (int o = 0; o < nouter; ++o) {
            while (cnt_o > 0 && nt->n[order[cnt_o - 1]] <= o * OB) --cnt_o;
            if (cnt_o == 0) break;
            const int slot = o & 1;
            if (s2 != s && cpass[slot]) CU_TRY(cudaStreamWaitEvent(s, nt->ev_cpass[slot], 0));
            if (legacy) {
               k_outer_begin<<<(cnt_o + 127) / 128, 128, 0, s>>>(T, d_fr, cnt_o, OB);
               ++launches;
            }
            int cnt_s = cnt_o;
            bool diag_done = false;      // the previous block's panel update already factorized this block's diagonal
            for (int ib = 0; ib < OB / IB; ++ib) {
               while (cnt_s > 0 && nt->n[order[cnt_s - 1]] <= o * OB + ib * IB) --cnt_s;
               if (cnt_s == 0) break;
               TileBatch rb{d_fr, d_row, cnt_s};
               TileBatch ub{d_fr, d_inn, cnt_s};
               // does another block column of this panel follow (for any front)?
               int cnt_next = cnt_s;
               while (cnt_next > 0 && nt->n[order[cnt_next - 1]] <= o * OB + (ib + 1) * IB) --cnt_next;
               const bool fuse_next = fuse_diag && !legacy && ib + 1 < OB / IB && cnt_next > 0;
               {
                  ProfScope ps(nt, KC_POTRF);
                  if (!diag_done) {
                     // the first block column of a panel also opens it (cnt_s == cnt_o then)
                     k_ldlt_diag32<<<cnt_s, 32, 0, s>>>(T, d_fr, d_diag, u, small, (ib == 0 && !legacy) ? OB : 0);
                     ++launches;
                  }
                  if (legacy) {
                     k_apply32<<<row_prefix[cnt_s], AP_THREADS, 0, s>>>(T, rb, d_diag, u, small);
                     k_finish32<<<row_prefix[cnt_s], AP_THREADS, 0, s>>>(T, rb, d_diag, small);
                     k_swap_failed<<<cnt_s, SW_THREADS, 0, s>>>(T, d_fr, d_diag);
                     launches += 3;
                  } else {
                     k_block_column32<<<row_prefix[cnt_s], AP_THREADS, 0, s>>>(T, rb, d_diag, u, small);
                     ++launches;
                  }
               }
               {
                  // rest of the panel, including the columns that just failed (CTAs of fronts
                  // with nothing left in the panel exit at once); fused: + the next diagonal block
                  ProfScope ps(nt, KC_TRSM);
                  if (fuse_next)
                     k_gemm_diag3<<<gemm_grid(3, inn_prefix[cnt_s]), GT_THREADS, GT_SMEM_BYTES, s>>>(T, ub, d_diag, u, small);
                  else
                     k_gemm_batched<<<gemm_grid(3, inn_prefix[cnt_s]), GT_THREADS, GT_SMEM_BYTES, s>>>(T, ub, 3, 0, IB, nullptr, 0, 0, 1);
                  ++launches;
                  diag_done = fuse_next;
               }
            }
            const bool upd = nt->n[order[0]] > (o + 1) * OB || o > 0;
            const bool split = upd && s2 != s;      // look-ahead: critical tile columns here, the rest on s2
            if (upd) {
               // columns behind the panel exist (more candidates, or failed columns of earlier panels)
               if (split && bulk) CU_TRY(cudaStreamWaitEvent(s, nt->ev_bulk, 0));      // previous panel's bulk part
               TileBatch ub{d_fr, d_upd, cnt_o};
               ProfScope ps(nt, KC_UPDATE);
               k_gemm_batched<<<gemm_grid(5, upd_prefix[cnt_o]), GT_THREADS, GT_SMEM_BYTES, s>>>(T, ub, 5, split ? 1 : 0, IB, nullptr, 0, 0, 1);
               ++launches;
            }
            k_outer_end<<<cnt_o, SW_THREADS, 0, s>>>(T, d_fr, slot);
            ++launches;
            if (!legacy && (split || con_prefix[cnt_o] > 0)) {
               if (s2 != s) {
                  CU_TRY(cudaEventRecord(nt->ev_panel, s));
                  CU_TRY(cudaStreamWaitEvent(s2, nt->ev_panel, 0));
               }
               if (split) {
                  TileBatch ub{d_fr, d_upd, cnt_o};
                  ProfScope ps(nt, KC_UPDATE, s2);
                  k_gemm_batched<<<gemm_grid(7, upd_prefix[cnt_o]), GT_THREADS, bulk_smem, s2>>>(T, ub, 7, slot, IB, nullptr, 0, 0, 1);
                  ++launches;
                  CU_TRY(cudaEventRecord(nt->ev_bulk, s2));
                  bulk = true;
               }
               if (con_prefix[cnt_o] > 0) {
                  TileBatch cb{d_fr, d_con, cnt_o};
                  ProfScope ps(nt, KC_CONTRIB, s2);
                  k_gemm_batched<<<gemm_grid(6, con_prefix[cnt_o]), GT_THREADS, bulk_smem, s2>>>(T, cb, 6, slot, IB, nullptr, 0, 0, 1);
                  ++launches;
               }
               if (s2 != s) {
                  CU_TRY(cudaEventRecord(nt->ev_cpass[slot], s2));
                  cpass[slot] = true;
               }
            }
         }
         if (bulk) CU_TRY(cudaStreamWaitEvent(s, nt->ev_bulk, 0));
         for (int slot = 0; slot < 2; ++slot)
            if (s2 != s && cpass[slot]) CU_TRY(cudaStreamWaitEvent(s, nt->ev_cpass[slot], 0));
         // ---- second pass (TPP) on failed columns, contribution blocks, statistics ----
         {
            ProfScope ps(nt, KC_ZERO);
            // G CTAs per front while the whole launch stays resident (the group barrier spins);
            // SYLVER_B200_TPP_G=1 selects the single-CTA kernel
            const char* tge = getenv("SYLVER_B200_TPP_G");
            const int gmax = tge ? std::max(1, std::min(TPP_GMAX, atoi(tge))) : TPP_GMAX;
            const int G = shared_device ? 1 : std::max(1, std::min(gmax, 128 / std::max(cnt, 1)));
            if (G > 1)
               k_tpp_multi<<<cnt * G, TPP_THREADS, 0, s>>>(T, d_fr, d_diag, u, small, tpp_everywhere ? 0 : 1, G);
            else
               k_tpp<<<cnt, TPP_THREADS, 0, s>>>(T, d_fr, u, small, tpp_everywhere ? 0 : 1);
            ++launches;
         }
         if (con_prefix[cnt] > 0) {
            TileBatch cb{d_fr, d_con, cnt};
            ProfScope ps(nt, KC_CONTRIB);
            k_gemm_batched<<<gemm_grid(4, con_prefix[cnt]), GT_THREADS, GT_SMEM_BYTES, s>>>(T, cb, 4, legacy ? 1 : 0, IB, nullptr, 0, 0, 1);
            ++launches;
         }
         assemble(1);
         k_front_stats<<<(cnt + 3) / 4, 128, 0, s>>>(T, d_fr, cnt, nt->d_stats, nt->d_lvl_out, d_slot);
         ++launches;
      }
      // ---- every rank learns the eliminated counts of the whole level ----
      if (multi && comm_allreduce_max_int(nt->d_lvl_out, (size_t)gcnt, s)) throw CudaFailure{-52};
      k_publish_nelim<<<(gcnt + 255) / 256, 256, 0, s>>>(st->d_level_nodes + gfirst, nt->d_lvl_out, gcnt, nt->d_nelim,
                                                         nt->d_state);
      ++launches;
      CU_TRY(cudaMemcpyAsync(nt->h_lvl_out, nt->d_lvl_out, gcnt * sizeof(int), cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      CU_TRY(cudaGetLastError());
      for (int i = 0; i < gcnt; ++i) nt->nelim[st->level_nodes[gfirst + i]] = nt->h_lvl_out[i];
      if (multi) exchange_level_indef(nt, l);
   }
   if (multi && comm_allreduce_sum_int(nt->d_stats, 8, s)) throw CudaFailure{-52};
   CU_TRY(cudaEventRecord(nt->ev1, s));
   int hs[8];
   CU_TRY(cudaMemcpyAsync(hs, nt->d_stats, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
   CU_TRY(cudaStreamSynchronize(s));
   float ms = 0;
   CU_TRY(cudaEventElapsedTime(&ms, nt->ev0, nt->ev1));
   nt->t_device = ms * 1e-3;
   nt->launches = launches;
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
   stats->num_delay = hs[0];
   stats->num_neg = hs[1];
   stats->num_two = hs[2];
   stats->num_zero = hs[3];
   stats->not_first_pass = hs[4];
   stats->not_second_pass = hs[5];
   if (nt->opt.pivot_method == 3) {
      // TPP is the first pass then (src/factor_failed.hxx:121-126)
      stats->not_first_pass = hs[5];
      stats->not_second_pass = 0;
   }
   stats->maxfront = maxfront;
   // roots cannot delay: what TPP leaves at a root are exact zero columns (counted above when
   // they were eliminated as zero pivots) -- anything else means a singular matrix and !action
   if (!nt->opt.action && hs[3] > 0) stats->flag = SYLVER_ERROR_SINGULAR;
   (void)N;
}

}  // namespace sylver_b200
