// Device kernels of the B200 multifrontal numeric factorization (sm_100a).
//
// Layout conventions (see DESIGN.md "Data layout in HBM"):
//   * Every front f owns an m x n column-major panel `L + loff[f]` with leading
//     dimension ldl[f] (multiple of 4 doubles, so every column starts 32 B
//     aligned and one element past an odd m is addressable padding).
//   * Its generated element (contribution block) is a dense (m-n) x (m-n)
//     lower triangle at `C + coff[f]`, leading dimension ldc[f].
//   * Per-edge assembly maps cmap (child contribution row -> parent local row)
//     are device resident and were built at symbolic-tree creation.
//
// Reference functions each kernel replaces are cited at the kernel.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sylver_b200 {

// Progress of the APTP factorization of one front (indefinite path only).
struct FrontState {
   int p0;       // columns eliminated so far: the contiguous prefix [0, p0)
   int na;       // end of the active candidates; [na, n) failed their block and wait for TPP
   int npass;    // columns of the current block that pass the a-posteriori test (atomicMin)
   int kbeg;     // pivots the next trailing update applies: [kbeg, kbeg + klen)
   int klen;
   int wb;       // width of the current block
   int nelim1;   // eliminated by the first (APTP) pass
   int nelim;    // eliminated in total; n - nelim columns are delayed to the parent
   // two-level blocking: the candidates are processed in outer panels of up to 256 columns;
   // inside a panel the 32-wide block columns update only the panel, the rest of the front
   // gets one rank-(p0 - obeg) update per panel
   int obeg;     // first column eliminated by the current outer panel
   int oend;     // end of the current outer panel (<= na)
   int pa;       // end of the panel's active candidates; [pa, oend) failed inside this panel
   int arrive;   // CTAs of the block-column kernel that finished the apply phase / the finish phase
   int done;
   int pad;
   // pivots [cb[s], ce[s]) of a completed outer panel, for its contribution-block pass (which runs
   // on the second stream while the next panel is being factorized; s = panel parity)
   int cb[2], ce[2];
   // look-ahead: the panel's end and whether its failed columns had to be swapped behind the
   // active candidates (then its whole trailing update ran on the critical path)
   int co[2], cf[2];
};
static_assert(sizeof(FrontState) == 88, "FrontState layout");

// Device view of the assembly tree (structure-of-arrays, one entry per front).
struct DevTree {
   const int* m;         // rows of the front (nrow + ndelay_in)
   const int* n;         // fully-summed columns (ncol + ndelay_in)
   const int* ldl;       // leading dimension of the L panel
   const int* ldc;       // leading dimension of the contribution block
   const long* loff;     // offset (doubles) of the L panel in the factor arena
   const long* coff;     // offset (doubles) of the contribution block
   const long* cmapoff;  // offset into cmap of this front's contribution rows
   const int* parent;    // parent front (nnodes = virtual root)
   const int* nchild;    // number of children
   const int* cmap;      // concatenated child->parent row maps (0-based)
   double* L;            // factor arena
   double* C;            // contribution arena
   // ---- indefinite path only ----
   const int* ncol0;     // fully-summed columns before delays (n - ncol0 = ndelay_in)
   const long* woff;     // offset of the front's W = L*D scratch panel (same shape as L)
   const long* doff;     // offset of D^-1 (2 doubles per column)
   const long* permoff;  // offset of the front's pivot permutation (n ints, 1-based variables)
   double* W;
   double* D;
   int* perm;
   FrontState* state;
   // ---- extend-add fused into the contribution-block epilogue (both paths) ----
   // Up to two children per front (the ones with the largest generated elements) are not
   // scattered into the parent's contribution block: the DMMA kernel that writes the block
   // gathers them instead.  pinv + pinvoff[2f+s] maps parent contribution row q to the row of
   // child fchild[2f+s]'s block that lands there (-1: none).  Static, built at analyse.
   const int* fchild;    // [2*nnodes], -1 = unused slot; nullptr disables the fusion
   const long* pinvoff;  // [2*nnodes]
   const int* pinv;
   // ---- fronts split over a rank group (multi-GPU, top of the tree) ----
   // splitP[f] > 1: the block columns (width 128) of front f are dealt round-robin to the
   // splitP[f] members of its group and this rank is member splitQ[f]; the extend-add only
   // touches the columns this rank owns.  nullptr: no front is split.
   const int* splitP;
   const int* splitQ;
};

// Does member q of a P-way split own column `col` of a front with n fully-summed columns?
// L panel: block column col/128.  Contribution block: tile column of the DMMA tile grid, which
// starts at the even column n & ~1 (k_gemm_batched mode 1).
__device__ __forceinline__ bool split_owns(int P, int q, int n, int col) {
   const int blk = (col < n) ? (col >> 7) : ((col - (n & ~1)) >> 7);
   return blk % P == q;
}

// ---------------------------------------------------------------------------
// PTX helpers: FP64 tensor-core MMA, mbarrier, TMA bulk copy
// ---------------------------------------------------------------------------
// D(8x8) += A(8x4) * B(4x8); fragment layout (lane = 4*g + t):
//   a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].   SASS: DMMA.8x8x4
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c0), "+d"(c1)
                : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
   return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier
// (SASS: UBLKCP).  dst/src 16 B aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
   asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
         smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------
// Tiled DMMA  C(128x64) (+)= -A(128xK) * B(64xK)^T   ("NT" rank-K update)
//
// Replaces update_block / update_contrib_block / form_contrib
// (reference src/kernels/factor.hxx:149-263, src/kernels/factor_indef.hxx:23-49,
// 201-342; GPU variants src/StarPU/cuda/kernels.hxx:45-182) and, with the
// inverted diagonal block as B, solve_block (src/kernels/factor.hxx:95-138).
//
// The host work lists count 128 x 128 tiles of a lower-triangular tile grid; every tile is
// computed by TWO CTAs (one per 64-column half, adjacent block indices so that they share the
// A panel in L2) and two CTAs are resident per SM (2 x 100 KB of shared memory, 2 x 8 warps at
// <= 128 registers): while one CTA reads and writes its destination tile (the read-modify-write
// epilogue of a rank-256 update is a fifth of the tile's life) the other one keeps the FP64
// tensor pipe busy, and the tail of a launch drains at half-tile granularity.
//
// Both operands are column-major panels (row index contiguous), staged into shared memory
// k-column by k-column with TMA bulk copies (cp.async.bulk, one per k-column and operand,
// into padded conflict-free rows) through a 4-stage mbarrier pipeline.  There is no producer
// warp: warp 0 issues the copies of k-block kb + 2 before it consumes block kb (a 17th warp
// would cap the kernel at 96 registers).  8 warps (4 x 2, 32 x 32 each) issue DMMA.8x8x4 with
// the front COLUMN as the MMA row index, so that a lane ends up with two vertically adjacent
// entries of the column-major destination: the epilogue reads and writes the destination
// straight from the accumulator registers with 16-byte accesses in full 64-byte segments.
// ---------------------------------------------------------------------------
constexpr int GT_BM = 128;             // tile rows
constexpr int GT_BN = 128;             // width of a tile-grid column (two CTAs of GT_HN columns)
constexpr int GT_HN = 64;              // columns per CTA
constexpr int GT_KT = 16;              // k-columns per stage
constexpr int GT_LDA = 132;            // padded smem row strides (mod 16 == 4: conflict free)
constexpr int GT_LDB = 68;
constexpr int GT_STAGES = 4;
constexpr int GT_LOOKAHEAD = 2;        // k-blocks in flight ahead of the one being consumed
constexpr int GT_LOOKAHEAD_TCOLS = 3;  // APTP look-ahead: tile columns of a panel update that stay on the critical path
constexpr int GT_WARPS = 8;            // consumer warps: 4 (rows) x 2 (columns), 32 x 32 each
#ifndef GT_PRODUCER_WARP
#define GT_PRODUCER_WARP 1             // 1: a ninth warp stages the operands (96-register cap); 0: the eight warps share the issue
#endif
constexpr int GT_THREADS = (GT_WARPS + GT_PRODUCER_WARP) * 32;
constexpr size_t GT_STAGE_DOUBLES = (size_t)GT_KT * (GT_LDA + GT_LDB);
constexpr size_t GT_SMEM_BYTES = GT_STAGES * GT_STAGE_DOUBLES * sizeof(double) + 2 * GT_STAGES * sizeof(uint64_t);
static_assert(2 * (GT_SMEM_BYTES + 1024) <= 232448, "two CTAs per SM");

// CTAs of a launch that covers `tiles` 128 x 128 tiles (mode 2 keeps one CTA per tile)
static inline int gemm_grid(int mode, int tiles) { return mode == 2 ? tiles : 2 * tiles; }

struct GemmTile {
   const double* A;   // row operand: first row of the tile, k = 0 column
   const double* B;   // column operand: first column of the CTA's half, k = 0 column
   int lda, ldb;      // column strides (doubles)
   int arows, brows;  // valid rows (<= 128) / columns (<= 64) of the tile
   int K;             // depth
};

struct GemmPipe {
   double* smem;
   uint64_t* full;
   uint64_t* empty;
   int gk;            // k-blocks that went through the pipeline so far (phase bookkeeping)
};

__device__ __forceinline__ void gemm_pipe_init(GemmPipe& p, double* smem) {
   p.smem = smem;
   p.full = reinterpret_cast<uint64_t*>(smem + GT_STAGES * GT_STAGE_DOUBLES);
   p.empty = p.full + GT_STAGES;
   p.gk = 0;
   if (threadIdx.x == 0) {
      for (int s = 0; s < GT_STAGES; ++s) {
         mbar_init(&p.full[s], GT_PRODUCER_WARP ? 1 : GT_WARPS);
         mbar_init(&p.empty[s], GT_WARPS);
      }
      mbar_fence_init();
   }
   __syncthreads();
}

// Stages k-block kb of the tile.  Producer-warp build: called by the ninth warp, lane l < 16
// copies A column l, lane 16 + l B column l.  Shared-issue build: called by every consumer warp,
// warp w copies k-columns 2w and 2w + 1 (lanes 0/1 the A columns, lanes 2/3 the B columns).
__device__ __forceinline__ void gemm_issue_block(const GemmPipe& p, const GemmTile& t, int kb, bool share, int warp,
                                                 int lane) {
   const int G = p.gk + kb;
   const int s = G % GT_STAGES;
   if (G >= GT_STAGES) mbar_wait(&p.empty[s], ((G / GT_STAGES) - 1) & 1);
   const int k0 = kb * GT_KT;
   const int kv = min(GT_KT, t.K - k0);           // valid k-columns in this stage
   double* As = p.smem + s * GT_STAGE_DOUBLES;
   double* Bs = As + GT_KT * GT_LDA;
   const uint32_t abytes = (uint32_t)(((t.arows + 1) & ~1) * sizeof(double));
   const uint32_t bbytes = (uint32_t)(((t.brows + 1) & ~1) * sizeof(double));
   const int kz = (kv + 3) & ~3;
#if GT_PRODUCER_WARP
   const int kc = lane & 15;
   const bool isB = lane >= 16;
   const bool mine = lane < 32;
   if (kc >= kv && kc < kz) {
      // zero-fill the tail of a partial k-group (generic proxy; ordered by the arrive below)
      if (!isB) {
         for (int r = 0; r < GT_BM; ++r) As[kc * GT_LDA + r] = 0.0;
      } else if (!share) {
         for (int r = 0; r < GT_HN; ++r) Bs[kc * GT_LDB + r] = 0.0;
      }
   }
   __syncwarp();
   if (lane == 0) mbar_expect_tx(&p.full[s], kv * (abytes + (share ? 0u : bbytes)));
   (void)warp;
#else
   const int kc = 2 * warp + (lane & 1);
   const bool isB = (lane & 2) != 0;
   const bool mine = lane < 4;
   if (kv < kz && 2 * warp + 1 >= kv && 2 * warp < kz) {
      for (int c = max(kv, 2 * warp); c < min(kz, 2 * warp + 2); ++c) {
         for (int r = lane; r < GT_BM; r += 32) As[c * GT_LDA + r] = 0.0;
         if (!share)
            for (int r = lane; r < GT_HN; r += 32) Bs[c * GT_LDB + r] = 0.0;
      }
   }
   __syncwarp();
   if (lane == 0) {
      // every warp arrives (its zero fill is ordered before the consumers' reads); warp 0 posts
      // the byte count of the whole stage -- copies of other warps may complete first, the
      // transaction count then goes negative for a moment
      if (warp == 0) mbar_expect_tx(&p.full[s], kv * (abytes + (share ? 0u : bbytes)));
      else mbar_arrive(&p.full[s]);
   }
#endif
   __syncwarp();
   if (mine && kc < kv) {
      if (!isB)
         tma_bulk_g2s(As + kc * GT_LDA, t.A + (size_t)(k0 + kc) * t.lda, abytes, &p.full[s]);
      else if (!share)
         tma_bulk_g2s(Bs + kc * GT_LDB, t.B + (size_t)(k0 + kc) * t.ldb, bbytes, &p.full[s]);
   }
}

// Accumulates the product tile: acc[ic][jr][h] = sum_k B[col][k] A[row][k] with
//   col = wn*32 + ic*8 + g,  row = wm*32 + jr*8 + 2*tq + h      (lane = 4*g + tq).
// `active` warps compute; the others only keep the pipeline moving.
__device__ __forceinline__ void gemm_tile_mainloop(GemmPipe& p, const GemmTile& t, bool active, double (&acc)[4][4][2]) {
   const int warp = threadIdx.x >> 5;
   const int lane = threadIdx.x & 31;
   const int nk = (t.K + GT_KT - 1) / GT_KT;
   const bool share = (t.A == t.B) && (t.lda == t.ldb);   // diagonal tile: the columns are rows of A
#if GT_PRODUCER_WARP
   if (warp == GT_WARPS) {
      // producer warp: runs ahead of the consumers by up to GT_STAGES k-blocks
      for (int kb = 0; kb < nk; ++kb) gemm_issue_block(p, t, kb, share, warp, lane);
      p.gk += nk;
      return;
   }
#endif
   const int wm = warp & 3, wn = warp >> 2;
   const int g = lane >> 2, tq = lane & 3;
#pragma unroll
   for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#if !GT_PRODUCER_WARP
   for (int kb = 0; kb < min(nk, GT_LOOKAHEAD); ++kb) gemm_issue_block(p, t, kb, share, warp, lane);
#endif
   for (int kb = 0; kb < nk; ++kb) {
#if !GT_PRODUCER_WARP
      if (kb + GT_LOOKAHEAD < nk) gemm_issue_block(p, t, kb + GT_LOOKAHEAD, share, warp, lane);
#endif
      const int G = p.gk + kb;
      const int s = G % GT_STAGES;
      mbar_wait(&p.full[s], (G / GT_STAGES) & 1);
      if (active) {
         const double* As = p.smem + s * GT_STAGE_DOUBLES;
         const double* Bs = share ? As : As + GT_KT * GT_LDA;
         const int ldb = share ? GT_LDA : GT_LDB;
         const int kv = min(GT_KT, t.K - kb * GT_KT);
         const int kg = (kv + 3) >> 2;
         const double* rp = As + tq * GT_LDA + wm * 32 + g;
         const double* cp = Bs + tq * ldb + wn * 32 + g;
#pragma unroll 1
         for (int k4 = 0; k4 < kg; ++k4) {
            double rf[4], cf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) cf[i] = cp[i * 8];
#pragma unroll
            for (int j = 0; j < 4; ++j) rf[j] = rp[j * 8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
               for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], cf[i], rf[j]);
            rp += 4 * GT_LDA;
            cp += 4 * ldb;
         }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&p.empty[s]);
   }
   p.gk += nk;
}

// Work descriptor of one batched launch: the participating fronts and an
// exclusive prefix sum of their tile counts.
struct TileBatch {
   const int* fronts;   // [cnt]
   const int* prefix;   // [cnt+1]
   int cnt;
};

__device__ __forceinline__ int find_front(const TileBatch& b, int item) {
   int lo = 0, hi = b.cnt;          // prefix[lo] <= item < prefix[hi]
   while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (b.prefix[mid] <= item) lo = mid; else hi = mid;
   }
   return lo;
}

// Destination of a tile: element (r, c) (absolute front indices) lives at dbase + c*ldd + r and is
// valid iff  c in [clo, chi),  r in [max(rmin, lower ? c : 0), m).
struct GemmDest {
   double* dbase;
   int ldd, clo, chi, rmin, m;
   bool lower;
   int op;              // 0: dst -= v   1: dst = gathered children - v   2: dst = v
   // fused extend-add (contribution tiles, op 1): up to two children whose generated elements
   // are gathered through parent-contribution-row -> child-row maps (-1: none, increasing)
   const double* gsrc[2];
   const int* gmap[2];
   int gld[2];
   int n;               // first contribution row/column of the front
};

// Epilogue of one warp's 32 x 32 block, straight from the accumulator registers.
__device__ __forceinline__ void gemm_tile_epilogue(const GemmDest& d, int i0, int j0, const double (&acc)[4][4][2]) {
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int wm = warp & 3, wn = warp >> 2;
   const int g = lane >> 2, tq = lane & 3;
   const int rb = i0 + wm * 32 + 2 * tq;        // + jr*8 + h
   const int cb = j0 + wn * 32 + g;             // + ic*8
   const int rlo = i0 + wm * 32, clo_w = j0 + wn * 32;
   // the whole 32 x 32 block valid?
   const bool interior = (rlo + 32 <= d.m) && (clo_w >= d.clo) && (clo_w + 32 <= d.chi) && (rlo >= d.rmin) &&
                         (!d.lower || rlo >= clo_w + 31);
   const bool vec = ((reinterpret_cast<uintptr_t>(d.dbase + rb) & 15) == 0) && ((d.ldd & 1) == 0);
   const bool fused = d.op == 1 && (d.gsrc[0] || d.gsrc[1]);
   int fir[2][4][2], fic[2][4];
   if (fused) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
#pragma unroll
         for (int jr = 0; jr < 4; ++jr)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
               const int r = rb + jr * 8 + h;
               fir[s][jr][h] = (d.gsrc[s] && r >= d.n && r < d.m) ? d.gmap[s][r - d.n] : -1;
            }
#pragma unroll
         for (int ic = 0; ic < 4; ++ic) {
            const int c = cb + ic * 8;
            fic[s][ic] = (d.gsrc[s] && c >= d.n && c < d.m) ? d.gmap[s][c - d.n] : -1;
         }
      }
   }
#pragma unroll
   for (int ic0 = 0; ic0 < 4; ic0 += 2) {
      double v[2][4][2];
      bool ok[2][4][2];
      double* colp[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
         const int ic = ic0 + u;
         const int c = cb + ic * 8;
         const bool cok = interior || (c >= d.clo && c < d.chi);
         colp[u] = d.dbase + (size_t)c * d.ldd;
#pragma unroll
         for (int jr = 0; jr < 4; ++jr)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
               const int r = rb + jr * 8 + h;
               ok[u][jr][h] = interior || (cok && r < d.m && r >= d.rmin && (!d.lower || r >= c));
               v[u][jr][h] = acc[ic][jr][h];
            }
      }
      if (fused) {
#pragma unroll
         for (int s = 0; s < 2; ++s) {
            if (!d.gsrc[s]) continue;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
               const int icc = fic[s][ic0 + u];
               const double* col = d.gsrc[s] + (size_t)max(icc, 0) * d.gld[s];
#pragma unroll
               for (int jr = 0; jr < 4; ++jr)
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                     const bool hit = icc >= 0 && fir[s][jr][h] >= icc;      // maps are increasing: row >= col
                     const double gv = hit ? col[fir[s][jr][h]] : 0.0;
                     v[u][jr][h] -= gv;
                  }
            }
         }
      }
      if (d.op == 0) {
         // read-modify-write: all loads of the batch first (8 x 16 B in flight per lane)
         double q[2][4][2];
#pragma unroll
         for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int jr = 0; jr < 4; ++jr) {
               double* ptr = colp[u] + rb + jr * 8;
               if (vec && ok[u][jr][0] && ok[u][jr][1]) {
                  const double2 x = *reinterpret_cast<const double2*>(ptr);
                  q[u][jr][0] = x.x; q[u][jr][1] = x.y;
               } else {
                  q[u][jr][0] = ok[u][jr][0] ? ptr[0] : 0.0;
                  q[u][jr][1] = ok[u][jr][1] ? ptr[1] : 0.0;
               }
            }
#pragma unroll
         for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int jr = 0; jr < 4; ++jr) {
               v[u][jr][0] = q[u][jr][0] - v[u][jr][0];
               v[u][jr][1] = q[u][jr][1] - v[u][jr][1];
            }
      } else if (d.op == 1) {
#pragma unroll
         for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int jr = 0; jr < 4; ++jr) { v[u][jr][0] = -v[u][jr][0]; v[u][jr][1] = -v[u][jr][1]; }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
         for (int jr = 0; jr < 4; ++jr) {
            double* ptr = colp[u] + rb + jr * 8;
            if (vec && ok[u][jr][0] && ok[u][jr][1]) {
               *reinterpret_cast<double2*>(ptr) = make_double2(v[u][jr][0], v[u][jr][1]);
            } else {
               if (ok[u][jr][0]) ptr[0] = v[u][jr][0];
               if (ok[u][jr][1]) ptr[1] = v[u][jr][1];
            }
         }
   }
}

// Pull what the epilogue of this warp will touch into L2 while the tensor cores work: the
// destination block (read-modify-write modes) or the children's entries it gathers.
__device__ __forceinline__ void gemm_tile_prefetch(const GemmDest& d, int i0, int j0) {
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int wm = warp & 3, wn = warp >> 2;
   const int r0 = i0 + wm * 32, c = j0 + wn * 32 + lane;
   if (d.op == 0) {
      if (c >= d.clo && c < d.chi) {
         const int ra = max(r0, max(d.rmin, d.lower ? c : 0)), rz = min(r0 + 32, d.m);
         if (ra < rz) {
            const char* p0 = reinterpret_cast<const char*>(d.dbase + (size_t)c * d.ldd + ra);
            const char* p1 = reinterpret_cast<const char*>(d.dbase + (size_t)c * d.ldd + rz - 1);
            for (const char* p = p0; p <= p1; p += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p1));
         }
      }
   } else if (d.op == 1 && (d.gsrc[0] || d.gsrc[1])) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
         if (!d.gsrc[s]) continue;
         const int r = r0 + lane;
         int lo = (r >= d.n && r < d.m) ? d.gmap[s][r - d.n] : -1;
         int hi = lo;
         if (lo < 0) lo = 0x7fffffff;
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
         }
         if (hi < 0) continue;
         const int ic = (c >= d.n && c < d.m) ? d.gmap[s][c - d.n] : -1;
         if (ic < 0) continue;
         const int a = max(lo, ic);
         if (a > hi) continue;
         const char* p0 = reinterpret_cast<const char*>(d.gsrc[s] + (size_t)ic * d.gld[s] + a);
         const char* p1 = reinterpret_cast<const char*>(d.gsrc[s] + (size_t)ic * d.gld[s] + hi);
         for (const char* p = p0; p <= p1; p += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
         asm volatile("prefetch.global.L2 [%0];" ::"l"(p1));
      }
   }
}

// mode 0: trailing update inside the L panel after block column [p0, p0+pw):
//         L[r][c] -= sum_k L[r][k] L[c][k],  c in [p0+pw, n), r in [c, m)
// mode 1: contribution block:  C[r-n][c-n] = (children) - sum_{k<n} L[r][k] L[c][k], n <= c <= r < m
// mode 2: panel solve with the inverted diagonal block W (pw x pw lower triangular, ld wld):
//         L[r][p0+c] = sum_{k<=c} L[r][p0+k] W[c][k],  r in [p0+pw, m)   (in place: one CTA per
//         128-row tile does the right half first, then the left half, which reads only its own
//         64 columns)
// nb is the block-column width; `step` the block column index.  Modes 0 and 1 enumerate the
// tile columns tstart, tstart + tstep, ... of the tile grid (look-ahead scheduling: the grid
// size limits a launch to the first tile column, tstart = 1 skips it; split fronts: the tile
// columns this rank owns).  Launch gemm_grid(mode, tiles) CTAs.
// Fused diagonal step (indefinite path, defined in kernels_indef.cuh): the CTA that updates the
// next 32 x 32 diagonal block of a front factorizes it right away (k_ldlt_diag32's work) instead
// of a launch of its own after the whole panel update.
struct DiagScratch;
struct FusedDiag {
   char* scratch;            // DiagScratch array of the launch (one per front of the batch)
   size_t stride;            // sizeof(DiagScratch)
   double u, small;
};
__device__ void ldlt_diag32_front(const DevTree& T, int f, DiagScratch* out, double u, double small, int begin_ob,
                                  double* S, double* dinv, int* lperm);

template <bool FUSED>
__device__ __forceinline__ void gemm_body(const DevTree& T, const TileBatch& batch, int mode, int step, int nb,
                                          const double* __restrict__ W, int wld, int tstart, int tstep,
                                          const FusedDiag& fd, double* smem) {
   const int half = (mode == 2) ? 0 : (blockIdx.x & 1);
   const int item = (mode == 2) ? blockIdx.x : (blockIdx.x >> 1);
   const int fi = find_front(batch, item);
   const int f = batch.fronts[fi];
   int local = item - batch.prefix[fi];
   const int m = T.m[f], n = T.n[f], ldl = T.ldl[f];
   double* Lf = T.L + T.loff[f];
   const int p0 = step * nb;
   const int pw = min(nb, n - p0);
   const int warp = threadIdx.x >> 5;
   const int wm = warp & 3, wn = warp >> 2;

   GemmTile t;
   GemmDest d;
   d.m = m; d.n = n;
   d.gsrc[0] = d.gsrc[1] = nullptr;
   int i0, j0;          // absolute front row / column of the CTA's tile origin
   int amax = GT_BM;    // rows of the CTA's tile (the right half of a diagonal tile has 64)
   double acc[4][4][2];
   if (mode == 2) {
      // tile origin rounded down to an even row so the TMA source stays 16 B aligned
      i0 = ((p0 + pw) & ~1) + local * GT_BM;
      t.A = Lf + (size_t)p0 * ldl + i0;
      t.lda = ldl;
      t.arows = min(GT_BM, m - i0);
      t.ldb = wld;
      d.dbase = Lf; d.ldd = ldl; d.rmin = p0 + pw; d.lower = false; d.op = 2;
      GemmPipe pipe;
      gemm_pipe_init(pipe, smem);
      const bool rows_ok = wm * 32 < t.arows;
      for (int hh = (pw > GT_HN ? 1 : 0); hh >= 0; --hh) {
         j0 = p0 + hh * GT_HN;
         t.B = W + (size_t)fi * wld * wld + hh * GT_HN;   // slot fi of this launch's inverse buffer
         t.brows = min(GT_HN, pw - hh * GT_HN);
         t.K = min(pw, (hh + 1) * GT_HN);                  // W is lower triangular: k <= c
         d.clo = j0; d.chi = j0 + t.brows;
         const bool active = warp < GT_WARPS && rows_ok && wn * 32 < t.brows;
         gemm_tile_mainloop(pipe, t, active, acc);
         if (active) gemm_tile_epilogue(d, i0, j0, acc);
      }
      return;
   }
   if (mode >= 3) {
      // ---- indefinite path: operands are W = L*D (A side) and L (B side); the extent of the
      // update is read from the device-resident front state (nothing here is known to the host)
      // mode 3: block column [kbeg, kbeg+klen) -> the rest of the current outer panel
      // mode 5: outer panel [obeg, p0) -> everything behind the panel
      // mode 4: all eliminated columns -> contribution block
      const FrontState st = T.state[f];
      const double* Wf = T.W + T.woff[f];
      // mode 6: pivots [cb, ce) of the completed outer panel of parity `step` -> contribution block
      // mode 4: pivots the second pass (TPP) eliminated, [nelim1, nelim) -> contribution block
      //         (step = 1: all pivots [0, nelim), the round-1 sequence without per-panel passes)
      // The pass that applies pivot 0 overwrites the block (and gathers the fused children); the
      // later ones read-modify-write it.  A front that eliminates nothing is written by mode 4.
      const bool contrib = mode == 4 || mode == 6;
      // Look-ahead split of the panel -> rest update (mode 5 with step = 1 / mode 7): the first
      // three tile columns hold the next panel's candidates and are updated on the critical path
      // (mode 5, live state); the rest (mode 7, from the snapshot k_outer_end published in slot
      // `step`) runs on the bulk stream beside the next panel's pivoting chain.  A panel whose
      // failed columns must be swapped with candidates from the far end (ns > 0) keeps its whole
      // update on the critical path: the swap reads rows and columns of the tail.
      const int sl = step & 1;
      const int kbeg = (mode == 3) ? st.kbeg : (mode == 5 ? st.obeg : (mode == 7 ? st.cb[sl] : (mode == 6 ? st.cb[sl] : (step ? 0 : st.nelim1))));
      t.K = (mode == 3) ? st.klen : (mode == 5 ? st.p0 - st.obeg : ((mode == 6 || mode == 7) ? st.ce[sl] - kbeg : st.nelim - kbeg));
      const bool first_pass = contrib && kbeg == 0;
      // FUSED (mode 3): the CTA of the front's first tile also factorizes the next diagonal block,
      // whether or not this block column passed any pivot
      if constexpr (FUSED) {
         if (local == 0 && half == 0 && !(t.K > 0 && st.oend > st.p0)) {
            if (threadIdx.x < 32)
               ldlt_diag32_front(T, f, reinterpret_cast<DiagScratch*>(fd.scratch + fi * fd.stride), fd.u, fd.small, 0, smem,
                                 smem + 32 * 33, reinterpret_cast<int*>(smem + 32 * 33 + 68));
            return;
         }
      }
      if (t.K == 0 && !(mode == 4 && first_pass)) return;
      const int first = (mode == 3) ? st.p0 : (mode == 5 ? st.oend : (mode == 7 ? st.co[sl] : n));     // first column/row of the updated region
      const int base = first & ~1;
      const int cend = (mode == 3) ? st.oend : ((mode == 5 || mode == 7) ? n : m);
      if (cend <= first) return;
      int tj_lo = 0, tj_hi = 0x7fffffff;
      if (mode == 5 && step == 1) {
         const int ns = min(st.oend - st.p0, st.na - st.oend);
         if (ns <= 0) tj_hi = GT_LOOKAHEAD_TCOLS;
      } else if (mode == 7) {
         if (st.cf[sl]) return;
         tj_lo = GT_LOOKAHEAD_TCOLS;
      }
      const int TR = (m - base + GT_BM - 1) / GT_BM;
      const int TC = (cend - base + GT_BN - 1) / GT_BN;
      int tj = 0;
      while (tj < TC && local >= TR - tj) { local -= TR - tj; ++tj; }
      if (tj >= TC || tj < tj_lo || tj >= tj_hi) return;
      const int ti = tj + local;
      i0 = base + ti * GT_BM;
      j0 = base + tj * GT_BN + half * GT_HN;
      if (ti == tj && half) { i0 += GT_HN; amax = GT_HN; }      // right half of a diagonal tile: only its lower 64 rows
      if (j0 >= cend || i0 >= m) return;
      t.A = Wf + (size_t)kbeg * ldl + i0;
      t.B = Lf + (size_t)kbeg * ldl + j0;
      t.lda = t.ldb = ldl;
      if (!contrib) {
         d.dbase = Lf; d.ldd = ldl; d.clo = first; d.chi = cend; d.rmin = 0; d.lower = true; d.op = 0;
      } else {
         const int ldc = T.ldc[f];
         d.dbase = T.C + T.coff[f] - (size_t)n * ldc - n; d.ldd = ldc; d.clo = n; d.chi = m; d.rmin = 0; d.lower = true;
         d.op = first_pass ? 1 : 0;      // the other children are extend-added afterwards (k_assemble_indef part 1)
      }
   } else {
      const int base = (mode == 0) ? (p0 + pw) : (n & ~1);
      const int cend = (mode == 0) ? n : m;
      const int TR = (m - base + GT_BM - 1) / GT_BM;
      const int TC = (cend - base + GT_BN - 1) / GT_BN;
      int tj = tstart;
      while (tj < TC && local >= TR - tj) { local -= TR - tj; tj += tstep; }
      if (tj >= TC) return;
      const int ti = tj + local;
      i0 = base + ti * GT_BM;
      j0 = base + tj * GT_BN + half * GT_HN;
      if (ti == tj && half) { i0 += GT_HN; amax = GT_HN; }
      if (j0 >= cend || i0 >= m) return;
      const int kbeg = (mode == 0) ? p0 : 0;
      t.K = (mode == 0) ? pw : n;
      t.A = Lf + (size_t)kbeg * ldl + i0;
      t.B = Lf + (size_t)kbeg * ldl + j0;
      t.lda = t.ldb = ldl;
      if (mode == 0) {
         d.dbase = Lf; d.ldd = ldl; d.clo = p0 + pw; d.chi = n; d.rmin = 0; d.lower = true; d.op = 0;
      } else {
         const int ldc = T.ldc[f];
         d.dbase = T.C + T.coff[f] - (size_t)n * ldc - n; d.ldd = ldc; d.clo = n; d.chi = m; d.rmin = 0; d.lower = true;
         d.op = 1;      // the other children are extend-added afterwards (k_assemble part 1)
      }
   }
   t.arows = min(amax, m - i0);
   t.brows = min(GT_HN, m - j0);
   if (d.op == 1 && T.fchild) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
         const int fc = T.fchild[2 * f + s];
         if (fc < 0) continue;
         d.gsrc[s] = T.C + T.coff[fc];
         d.gld[s] = T.ldc[fc];
         d.gmap[s] = T.pinv + T.pinvoff[2 * f + s];
      }
   }
   GemmPipe pipe;
   gemm_pipe_init(pipe, smem);
   // warps whose 32 x 32 block lies outside the tile or strictly above the diagonal do no math
   const bool active = warp < GT_WARPS && (wm * 32 < t.arows) && (wn * 32 < t.brows) && (i0 + wm * 32 + 31 >= j0 + wn * 32);
   if (active) gemm_tile_prefetch(d, i0, j0);
   gemm_tile_mainloop(pipe, t, active, acc);
   if (active) gemm_tile_epilogue(d, i0, j0, acc);
   if constexpr (FUSED) {
      // first tile of the front (ti = tj = 0, left half): it holds the next diagonal block
      // [p0, p0 + 32)^2 entirely; once all warps of this CTA have written their part, warp 0
      // factorizes it (the pipeline buffers are free by now)
      if (item == batch.prefix[fi] && half == 0) {
         __syncthreads();
         if (threadIdx.x < 32)
            ldlt_diag32_front(T, f, reinterpret_cast<DiagScratch*>(fd.scratch + fi * fd.stride), fd.u, fd.small, 0, smem,
                              smem + 32 * 33, reinterpret_cast<int*>(smem + 32 * 33 + 68));
      }
   }
}

static __global__ void __launch_bounds__(GT_THREADS, 2)
k_gemm_batched(DevTree T, TileBatch batch, int mode, int step, int nb, const double* __restrict__ W, int wld,
               int tstart, int tstep) {
   extern __shared__ __align__(128) double smem[];
   gemm_body<false>(T, batch, mode, step, nb, W, wld, tstart, tstep, FusedDiag{nullptr, 0, 0.0, 0.0}, smem);
}

// ---------------------------------------------------------------------------
// A -> front scatter.  Replaces init_a_block / init_node
// (reference src/kernels/assemble.hxx:162-214): dest = col*nrow + row (1-based),
// rows >= ncol shifted by ndelay_in, optional symmetric scaling.
// ---------------------------------------------------------------------------
static __global__ void k_scatter_a(DevTree T, long nent, const long* __restrict__ nlist,
                            const int* __restrict__ anode, const int* __restrict__ nrow0,
                            const int* __restrict__ ncol0, const double* __restrict__ aval,
                            const double* __restrict__ scaling, const int* __restrict__ rlist,
                            const long* __restrict__ rptr, const int* __restrict__ owner, int me) {
   for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < nent; e += (long)gridDim.x * blockDim.x) {
      const int f = anode[e];
      if (owner && owner[f] != me) continue;      // multi-GPU: only the fronts this rank factorizes
      const long src = nlist[2 * e] - 1;
      const long dest = nlist[2 * e + 1] - 1;
      const int nrow = nrow0[f];
      const int c = (int)(dest / nrow);
      int r = (int)(dest - (long)c * nrow);
      double v = aval[src];
      if (scaling) {
         const int* rl = rlist + (rptr[f] - 1);
         v *= scaling[rl[r] - 1] * scaling[rl[c] - 1];
      }
      const int ndelay = T.n[f] - ncol0[f];
      if (r >= ncol0[f]) r += ndelay;
      T.L[T.loff[f] + (size_t)c * T.ldl[f] + r] = v;
   }
}

// ---------------------------------------------------------------------------
// Extend-add of one child's contribution block into its parent front.
// Replaces assemble_block / assemble_contrib_block (reference
// src/kernels/assemble.hxx:230-286,343-517) with a single pass driven by the
// precomputed device map: entry (i,j) of the child's block goes to parent
// (cmap[i], cmap[j]) -- into the L panel when cmap[j] < n_parent, else into the
// parent's contribution block.  One launch handles the q-th child of every
// parent of a level, so contributions are summed in a fixed (reference) order.
// work item = (child, first column of a 32-column chunk); warp w takes columns
// w, w+8, ...; lanes stride over rows (coalesced reads of the child block).
// ---------------------------------------------------------------------------
// part 0: columns that land in the parent's fully-summed (L) panel -- before the parent is
//         factorized;  part 1: columns that land in the parent's contribution block -- after
//         the parent's own Schur complement has been written there (the reference's order:
//         assemble_contrib follows factor_front, src/NumericTreePosdef.hxx:276-321), so the
//         contribution block needs no zero fill and the DMMA epilogue never reads it.
static __global__ void __launch_bounds__(256) k_assemble(DevTree T, const int2* __restrict__ work, int part) {
   const int2 w = work[blockIdx.x];
   const int c = w.x;
   const int p = T.parent[c];
   const int k = T.m[c] - T.n[c];
   const int* cm = T.cmap + T.cmapoff[c];
   const int pn = T.n[p];
   const int jend = min(k, w.y + 32);
   // cm is increasing: the chunk is skipped as a whole when it lies in the other part
   if (part == 0 ? (cm[w.y] >= pn) : (cm[jend - 1] < pn)) return;
   const int sP = T.splitP ? T.splitP[p] : 1;
   const int sQ = sP > 1 ? T.splitQ[p] : 0;
   // the contribution part of a fused child is gathered by the parent's DMMA epilogue
   if (part == 1 && T.fchild && (T.fchild[2 * p] == c || T.fchild[2 * p + 1] == c)) return;
   const double* src = T.C + T.coff[c];
   const int ldcc = T.ldc[c];
   const int pldl = T.ldl[p], pldc = T.ldc[p];
   double* PL = T.L + T.loff[p];
   double* PC = T.C + T.coff[p];
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   for (int j = w.y + warp; j < jend; j += 8) {
      const int rj = cm[j];
      if ((rj < pn) != (part == 0)) continue;
      if (sP > 1 && !split_owns(sP, sQ, pn, rj)) continue;      // another member's column
      const double* s = src + (size_t)j * ldcc;
      double* dcol = (rj < pn) ? PL + (size_t)rj * pldl : PC + (size_t)(rj - pn) * pldc - pn;
      int i = j + lane;
      for (; i + 96 < k; i += 128) {      // 4 independent read-modify-writes in flight per lane
         const int r0 = cm[i], r1 = cm[i + 32], r2 = cm[i + 64], r3 = cm[i + 96];
         const double v0 = s[i], v1 = s[i + 32], v2 = s[i + 64], v3 = s[i + 96];
         const double d0 = dcol[r0], d1 = dcol[r1], d2 = dcol[r2], d3 = dcol[r3];
         dcol[r0] = d0 + v0; dcol[r1] = d1 + v1; dcol[r2] = d2 + v2; dcol[r3] = d3 + v3;
      }
      for (; i < k; i += 32) dcol[cm[i]] += s[i];
   }
}

// ---------------------------------------------------------------------------
// Cholesky of one diagonal block (pw <= PW) per CTA, plus its explicit inverse
// W = L_jj^{-1} (lower) used by the DMMA panel solve.  Replaces factorize_diag_block
// (reference src/kernels/factor.hxx:34-88; LAPACK dpotrf there).  A non-positive pivot
// records (front, column) in `fail`.
//
// Three template sizes (32/64/128) so that the thousands of small fronts of the lower tree
// levels do not pay for a 200 KB shared-memory footprint (several CTAs per SM there).
// ---------------------------------------------------------------------------
constexpr int PF_THREADS = 512;      // PW = 128 variant (threads = 4 * PW)
template <int PW>
struct PotrfCfg {
   static constexpr int LD = PW + 1;
   static constexpr int NT = 4 * PW;
   static constexpr int VPACK = PW * (PW + 1) / 2;
   static constexpr size_t SMEM = ((size_t)PW * LD + VPACK + PW) * sizeof(double);
};
constexpr size_t PF_SMEM_BYTES = PotrfCfg<128>::SMEM;
// packed column-major lower triangle: element (i,j), i >= j
template <int PW>
__device__ __forceinline__ int vpk(int i, int j) { return j * PW - (j * (j - 1)) / 2 + (i - j); }

template <int PW>
static __global__ void __launch_bounds__(4 * PW, 1)
k_potrf_inv(DevTree T, const int* __restrict__ fronts, int step, int nb, double* __restrict__ W, int wld,
            int* fail) {
   using Cfg = PotrfCfg<PW>;
   constexpr int LD = Cfg::LD, NT = Cfg::NT, NW = NT / 32;
   extern __shared__ __align__(16) double sm[];
   double* S = sm;                    // S[r + c*LD]
   double* V = sm + PW * LD;          // inverse, packed lower triangle
   double* rd = V + Cfg::VPACK;       // reciprocal diagonal
   const int f = fronts[blockIdx.x];
   const int n = T.n[f], ldl = T.ldl[f];
   const int p0 = step * nb;
   const int pw = min(nb, n - p0);
   double* A = T.L + T.loff[f] + (size_t)p0 * ldl + p0;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   for (int c = warp; c < pw; c += NW)
      for (int r = lane; r < pw; r += 32) S[r + c * LD] = (r >= c) ? A[(size_t)c * ldl + r] : 0.0;
   __syncthreads();
   // right-looking Cholesky, 2 barriers per column; warp w updates columns k+1+w, k+1+w+NW, ...
   // (a blocked variant with a register-resident 32x32 warp factorization was measured 1.5x
   // SLOWER: the single-warp sections are latency bound at ~12 cycles per instruction)
   for (int k = 0; k < pw; ++k) {
      const double akk = S[k + k * LD];
      const bool ok = akk > 0.0;
      if (!ok && tid == 0) {
         atomicExch(&fail[0], 1);
         atomicMin(&fail[1], f);
      }
      const double rinv = ok ? rsqrt(akk) : 1.0;
      const double d = ok ? akk * rinv : 1.0;
      __syncthreads();
      for (int r = k + tid; r < pw; r += NT) S[r + k * LD] = (r == k) ? d : S[r + k * LD] * rinv;
      __syncthreads();
      for (int c = k + 1 + warp; c < pw; c += NW) {
         const double lck = S[c + k * LD];
         for (int r = c + lane; r < pw; r += 32) S[r + c * LD] -= S[r + k * LD] * lck;
      }
   }
   __syncthreads();
   for (int c = warp; c < pw; c += NW)
      for (int r = c + lane; r < pw; r += 32) A[(size_t)c * ldl + r] = S[r + c * LD];
   for (int i = tid; i < pw; i += NT) rd[i] = 1.0 / S[i + i * LD];
   __syncthreads();
   // Inverse by forward substitution: 4 lanes cooperate on one column j of V (L v = e_j),
   // splitting each dot product; no block barrier is needed inside a column.
   {
      const int q = tid & 3;
      for (int j = tid >> 2; j < PW; j += NT / 4) {
         if (j >= pw) continue;     // whole quad skips together (same j)
         if (q == 0) V[vpk<PW>(j, j)] = rd[j];
         __syncwarp(0xFu << (lane & ~3));
         for (int i = j + 1; i < pw; ++i) {
            double s = 0.0;
            const double* vj = V + vpk<PW>(j, j) - j;      // vj[kk] = V(kk, j)
            for (int kk = j + q; kk < i; kk += 4) s += S[i + kk * LD] * vj[kk];
            s += __shfl_xor_sync(0xFu << (lane & ~3), s, 1);
            s += __shfl_xor_sync(0xFu << (lane & ~3), s, 2);
            if (q == 0) V[vpk<PW>(i, j)] = -s * rd[i];
            __syncwarp(0xFu << (lane & ~3));
         }
      }
   }
   __syncthreads();
   double* Wf = W + (size_t)blockIdx.x * wld * wld;
   for (int c = warp; c < wld; c += NW)
      for (int r = lane; r < wld; r += 32)
         Wf[r + (size_t)c * wld] = (r < pw && c < pw && r >= c) ? V[vpk<PW>(r, c)] : 0.0;
}

// ---------------------------------------------------------------------------
// Register-resident 128 x 128 Cholesky + inverse (the block-column critical path of the large
// fronts).  Same contract as k_potrf_inv<128>.  512 threads; thread (lane, warp) keeps the
// elements (r = lane + 32 i, c = warp + 16 j), i < 4, j < 8 -- a 2-D cyclic layout, so the
// shrinking trailing matrix stays spread over all warps.  One pass over k computes BOTH
// factors by outer products with a single barrier per column:
//     trailing matrix  A(r,c) -= L(r,k) L(c,k)             (c > k)
//     inverse          T(r,c) -= L(r,k) V(k,c), V(k,:) = T(k,:) / L(k,k)   (c < k; T starts as I)
// Column c of A is dead once it has been eliminated and column c of T is born at that very
// step, so both live in the same 32 registers.  Shared memory only carries the broadcast of
// column k of A and row k of T (double buffered, 4 KB).
// ---------------------------------------------------------------------------
constexpr int PR_THREADS = 512;
static __global__ void __launch_bounds__(PR_THREADS, 1)
k_potrf_inv_reg(DevTree T, const int* __restrict__ fronts, int step, int nb, double* __restrict__ W, int wld,
                int* fail) {
   __shared__ double cbuf[2][128];
   __shared__ double vbuf[2][128];
   const int f = fronts[blockIdx.x];
   const int n = T.n[f], ldl = T.ldl[f];
   const int p0 = step * nb;
   const int pw = min(nb, n - p0);
   double* A = T.L + T.loff[f] + (size_t)p0 * ldl + p0;
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   double x[4][8];
#pragma unroll
   for (int j = 0; j < 8; ++j) {
      const int c = warp + 16 * j;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
         const int r = lane + 32 * i;
         x[i][j] = (r < pw && c < pw && r >= c) ? A[(size_t)c * ldl + r] : (r == c ? 1.0 : 0.0);
      }
   }
#pragma unroll
   for (int jk = 0; jk < 8; ++jk) {
      const int ik = jk >> 1;                    // k = 16 jk + kk lies in row block ik
#pragma unroll 1
      for (int kk = 0; kk < 16; ++kk) {
         const int k = 16 * jk + kk;
         if (k >= pw) break;
         const int lk = k & 31;
         double* cb = cbuf[k & 1];
         double* vb = vbuf[k & 1];
         if (warp == kk) {
#pragma unroll
            for (int i = 0; i < 4; ++i) cb[lane + 32 * i] = x[i][jk];
         }
         if (lane == lk) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
               if (j < jk || (j == jk && warp < kk)) vb[warp + 16 * j] = x[ik][j];
         }
         __syncthreads();
         const double akk = cb[k];
         const bool ok = akk > 0.0;
         if (!ok && threadIdx.x == 0) {
            atomicExch(&fail[0], 1);
            atomicMin(&fail[1], f);
         }
         const double rinv = ok ? rsqrt(akk) : 1.0;
         double lr[4];
#pragma unroll
         for (int i = 0; i < 4; ++i) {
            const int r = lane + 32 * i;
            lr[i] = (i >= ik && r > k) ? cb[r] * rinv : 0.0;
         }
#pragma unroll
         for (int j = 0; j < 8; ++j) {
            const int c = warp + 16 * j;
            if (j > jk || (j == jk && warp > kk)) {
               // trailing matrix column c > k
               const double lc = cb[c] * rinv;
#pragma unroll
               for (int i = 0; i < 4; ++i)
                  if (i >= ik) x[i][j] -= lr[i] * lc;
            } else if (j < jk || warp < kk) {
               // inverse column c < k
               const double vc = vb[c] * rinv;
#pragma unroll
               for (int i = 0; i < 4; ++i)
                  if (i >= ik) x[i][j] -= lr[i] * vc;
               if (lane == lk) x[ik][j] = vc;      // row k of the inverse is final
            } else {
               // c == k: column k of L is final (written out), column k of T is born
#pragma unroll
               for (int i = 0; i < 4; ++i) {
                  const int r = lane + 32 * i;
                  if (i >= ik) {
                     if (r >= k && r < pw) A[(size_t)k * ldl + r] = (r == k) ? akk * rinv : lr[i];
                     x[i][j] = (r == k) ? rinv : -lr[i] * rinv;
                  } else {
                     x[i][j] = 0.0;
                  }
               }
            }
         }
      }
   }
   double* Wf = W + (size_t)blockIdx.x * wld * wld;
#pragma unroll
   for (int j = 0; j < 8; ++j) {
      const int c = warp + 16 * j;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
         const int r = lane + 32 * i;
         if (r < wld && c < wld) Wf[r + (size_t)c * wld] = (r < pw && c < pw && r >= c) ? x[i][j] : 0.0;
      }
   }
}

// ---------------------------------------------------------------------------
// Blocked 128 x 128 Cholesky + inverse on one SM -- the block-column critical path of the large
// fronts (and of the multi-GPU panel chain).  Same contract as k_potrf_inv<128>.
//
// The column-at-a-time kernels above pay one CTA barrier (and a dependent rsqrt) per column,
// 128 of them in sequence.  Here the block is factorized in four 32-column steps:
//   (1) one warp holds the 32 x 32 diagonal block in registers (lane = row) and factorizes it
//       with shuffles -- no barrier inside the 32 columns;
//   (2) the rows below it are solved against L_jj^T, one thread per row, by substitution, while
//       another warp forms the 32 x 32 inverse V_jj (lane = column);
//   (3) the trailing blocks are updated with DMMA block products straight from shared memory.
// The off-diagonal blocks of the inverse follow block sub-diagonal by sub-diagonal,
//   V_ij = -V_ii (sum_{k=j}^{i-1} L_ik V_kj),  again as DMMA block products.
// Shared memory: S (128 x 132: L in the lower triangle, V^T in the strict upper triangle of the
// off-diagonal blocks), the four dense diagonal inverses, three product scratch blocks.
// ---------------------------------------------------------------------------
constexpr int PB_THREADS = 256;
constexpr int PB_LD = 132;       // == 4 (mod 16): DMMA fragment loads are conflict free
constexpr int PB_VLD = 36;
constexpr size_t PB_SMEM_BYTES = ((size_t)128 * PB_LD + 4 * 32 * PB_VLD + 3 * 32 * PB_VLD + 128) * sizeof(double);

// acc[ia][jb][h] += sum_{c < 32} X[ia*8 + g][c] * Y[jb*8 + 2*tq + h][c],  ia < NRA, jb < 4,
// X[a][c] = xs[a*xrs + c*xks],  Y[b][c] = ys[b*yrs + c*yks]     (lane = 4*g + tq)
template <int NRA>
__device__ __forceinline__ void warp_blockprod32(double (&acc)[NRA][4][2], const double* xs, int xrs, int xks,
                                                 const double* ys, int yrs, int yks) {
   const int lane = threadIdx.x & 31;
   const int g = lane >> 2, tq = lane & 3;
#pragma unroll
   for (int k4 = 0; k4 < 8; ++k4) {
      double a[NRA], b[4];
#pragma unroll
      for (int ia = 0; ia < NRA; ++ia) a[ia] = xs[(ia * 8 + g) * xrs + (k4 * 4 + tq) * xks];
#pragma unroll
      for (int jb = 0; jb < 4; ++jb) b[jb] = ys[(jb * 8 + g) * yrs + (k4 * 4 + tq) * yks];
#pragma unroll
      for (int ia = 0; ia < NRA; ++ia)
#pragma unroll
         for (int jb = 0; jb < 4; ++jb) dmma884(acc[ia][jb][0], acc[ia][jb][1], a[ia], b[jb]);
   }
}

static __global__ void __launch_bounds__(PB_THREADS, 1)
k_potrf_inv_blk(DevTree T, const int* __restrict__ fronts, int step, int nb, double* __restrict__ W, int wld,
                int* fail) {
   extern __shared__ __align__(16) double sm[];
   double* S = sm;                           // S[r*PB_LD + c]
   double* Vd = S + 128 * PB_LD;             // Vd[j][a*PB_VLD + b] = V_jj[a][b] (dense, zeros above the diagonal)
   double* Tb = Vd + 4 * 32 * PB_VLD;        // three 32 x 32 scratch blocks
   double* rinv = Tb + 3 * 32 * PB_VLD;      // 1 / L(r,r)
   const int f = fronts[blockIdx.x];
   const int n = T.n[f], ldl = T.ldl[f];
   const int p0 = step * nb;
   const int pw = min(nb, n - p0);
   double* A = T.L + T.loff[f] + (size_t)p0 * ldl + p0;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int g = lane >> 2, tq = lane & 3;
   const int nblk = (pw + 31) >> 5;
   // ---- load the lower triangle (coalesced down the columns); identity outside pw ----
   for (int c = warp; c < 32 * nblk; c += PB_THREADS / 32)
      for (int r = lane; r < 32 * nblk; r += 32) {
         double v = (r == c) ? 1.0 : 0.0;
         if (r < pw && c < pw && r >= c) v = A[(size_t)c * ldl + r];
         S[r * PB_LD + c] = v;
      }
   __syncthreads();
   for (int j = 0; j < nblk; ++j) {
      const int j0 = 32 * j;
      if (warp == 0) {
         // ---- (1) 32 x 32 Cholesky in registers: lane = row ----
         double x[32];
#pragma unroll
         for (int c = 0; c < 32; ++c) x[c] = S[(j0 + lane) * PB_LD + j0 + c];
         double rv = 1.0;
         bool bad = false;
#pragma unroll
         for (int k = 0; k < 32; ++k) {
            const double d = __shfl_sync(0xffffffffu, x[k], k);
            const bool ok = d > 0.0;
            bad |= !ok;
            const double ri = ok ? rsqrt(d) : 1.0;
            const double lk = (lane > k) ? x[k] * ri : (lane == k ? (ok ? d * ri : 1.0) : 0.0);
            x[k] = lk;
            if (lane == k) rv = ri;
#pragma unroll
            for (int c = k + 1; c < 32; ++c) {
               const double lc = __shfl_sync(0xffffffffu, lk, c);
               x[c] -= lk * lc;
            }
         }
         if (bad && lane == 0) {
            atomicExch(&fail[0], 1);
            atomicMin(&fail[1], f);
         }
#pragma unroll
         for (int c = 0; c < 32; ++c)
            if (c <= lane) S[(j0 + lane) * PB_LD + j0 + c] = x[c];
         rinv[j0 + lane] = rv;
      }
      __syncthreads();
      if (warp == 7) {
         // ---- (2b) V_jj = L_jj^{-1}: lane = column, rows in sequence (forward substitution) ----
         double v[32];
         const double* Lj = S + j0 * PB_LD + j0;
#pragma unroll
         for (int r = 0; r < 32; ++r) {
            double s0 = (r == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k < r; k += 2) {
               s0 -= Lj[r * PB_LD + k] * v[k];
               if (k + 1 < r) s1 -= Lj[r * PB_LD + k + 1] * v[k + 1];
            }
            v[r] = (s0 + s1) * rinv[j0 + r];
         }
#pragma unroll
         for (int r = 0; r < 32; ++r) Vd[j * 32 * PB_VLD + r * PB_VLD + lane] = v[r];
      } else if (tid < 32 * (nblk - 1 - j)) {
         // ---- (2a) rows below the diagonal block: w L_jj^T = a, one thread per row ----
         const int row = j0 + 32 + tid;
         double* Sr = S + row * PB_LD + j0;
         const double* Lj = S + j0 * PB_LD + j0;
         double w[32];
#pragma unroll
         for (int c = 0; c < 32; ++c) w[c] = Sr[c];
#pragma unroll
         for (int c = 0; c < 32; ++c) {
            double s0 = w[c], s1 = 0.0;
#pragma unroll
            for (int k = 0; k < c; k += 2) {
               s0 -= w[k] * Lj[c * PB_LD + k];
               if (k + 1 < c) s1 -= w[k + 1] * Lj[c * PB_LD + k + 1];
            }
            w[c] = (s0 + s1) * rinv[j0 + c];
         }
#pragma unroll
         for (int c = 0; c < 32; ++c) Sr[c] = w[c];
      }
      __syncthreads();
      // ---- (3) trailing update: S_ik -= L_ij L_kj^T for i >= k > j, one 32 x 32 block per warp ----
      {
         const int nrem = nblk - 1 - j;
         int t = warp, bi = 0, bk = 0;
         bool have = false;
         for (int i = 0; i < nrem && !have; ++i)
            for (int k = 0; k <= i; ++k) {
               if (t == 0) { bi = i; bk = k; have = true; break; }
               --t;
            }
         if (have) {
            const int i0 = j0 + 32 * (bi + 1), k0 = j0 + 32 * (bk + 1);
            double acc[4][4][2];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
               for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
            warp_blockprod32<4>(acc, S + i0 * PB_LD + j0, PB_LD, 1, S + k0 * PB_LD + j0, PB_LD, 1);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
               for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int h = 0; h < 2; ++h) S[(i0 + a * 8 + g) * PB_LD + k0 + b * 8 + 2 * tq + h] -= acc[a][b][h];
         }
      }
      __syncthreads();
   }
   // ---- L back to the panel (the update phase scribbled on the upper triangles: lower only) ----
   for (int c = warp; c < pw; c += PB_THREADS / 32)
      for (int r = c + lane; r < pw; r += 32) A[(size_t)c * ldl + r] = S[r * PB_LD + c];
   __syncthreads();
   // ---- off-diagonal blocks of the inverse, block sub-diagonal d: V_ij, i - j = d ----
   for (int d = 1; d < nblk; ++d) {
      const int ntask = nblk - d;                 // 3, 2, 1
      const int wpt = (d == 1) ? 2 : 4;           // warps per task (16 or 8 rows each)
      const int task = warp / wpt, part = warp % wpt;
      const int rows = 32 / wpt;                  // 16 or 8
      const bool have = task < ntask;
      const int bi = d + task, bj = task;
      const int i0 = 32 * bi, jj0 = 32 * bj;
      const int a0 = part * rows;
      double acc[2][4][2];
      if (have) {
#pragma unroll
         for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
         // T = sum_k L_ik V_kj   (X = L_ik rows a0.., Y[b][c] = V_kj[c][b])
         for (int k = bj; k < bi; ++k) {
            const double* xs = S + (i0 + a0) * PB_LD + 32 * k;
            const double* ys;
            int yrs, yks;
            if (k == bj) { ys = Vd + bj * 32 * PB_VLD; yrs = 1; yks = PB_VLD; }
            else { ys = S + jj0 * PB_LD + 32 * k; yrs = PB_LD; yks = 1; }      // V_kj^T lives at S[j-rows][k-cols]
            if (rows == 16) warp_blockprod32<2>(acc, xs, PB_LD, 1, ys, yrs, yks);
            else warp_blockprod32<1>(reinterpret_cast<double(&)[1][4][2]>(acc), xs, PB_LD, 1, ys, yrs, yks);
         }
         double* Tt = Tb + task * 32 * PB_VLD;
#pragma unroll
         for (int a = 0; a < 2; ++a)
            if (a * 8 < rows)
#pragma unroll
               for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int h = 0; h < 2; ++h) Tt[(a0 + a * 8 + g) * PB_VLD + b * 8 + 2 * tq + h] = acc[a][b][h];
      }
      __syncthreads();
      if (have) {
         // V_ij = -V_ii T   (X = Vd[i] rows a0.., Y[b][c] = T[c][b])
#pragma unroll
         for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
         const double* xs = Vd + bi * 32 * PB_VLD + a0 * PB_VLD;
         const double* ys = Tb + task * 32 * PB_VLD;
         if (rows == 16) warp_blockprod32<2>(acc, xs, PB_VLD, 1, ys, 1, PB_VLD);
         else warp_blockprod32<1>(reinterpret_cast<double(&)[1][4][2]>(acc), xs, PB_VLD, 1, ys, 1, PB_VLD);
         // transposed into the upper triangle: S[j0 + b][i0 + a] = V_ij[a][b]
#pragma unroll
         for (int a = 0; a < 2; ++a)
            if (a * 8 < rows)
#pragma unroll
               for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int h = 0; h < 2; ++h)
                     S[(jj0 + b * 8 + 2 * tq + h) * PB_LD + i0 + a0 + a * 8 + g] = -acc[a][b][h];
      }
      __syncthreads();
   }
   // ---- W = V (lower), zero elsewhere; coalesced down the columns ----
   double* Wf = W + (size_t)blockIdx.x * wld * wld;
   for (int c = warp; c < wld; c += PB_THREADS / 32)
      for (int r = lane; r < wld; r += 32) {
         double v = 0.0;
         if (r < pw && c < pw && r >= c) {
            if ((r >> 5) == (c >> 5)) v = Vd[(r >> 5) * 32 * PB_VLD + (r & 31) * PB_VLD + (c & 31)];
            else v = S[c * PB_LD + r];
         }
         Wf[r + (size_t)c * wld] = v;
      }
}

}  // namespace sylver_b200
