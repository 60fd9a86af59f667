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
   // two-level blocking: the candidates are processed in outer panels of up to 128 columns;
   // inside a panel the 32-wide block columns update only the panel, the rest of the front
   // gets one rank-(p0 - obeg) update per panel
   int obeg;     // first column eliminated by the current outer panel
   int oend;     // end of the current outer panel (<= na)
   int pa;       // end of the panel's active candidates; [pa, oend) failed inside this panel
   int pad;
};

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
// Tiled DMMA  C(128x128) (+)= -A(128xK) * B(128xK)^T   ("NT" rank-K update)
//
// Replaces update_block / update_contrib_block / form_contrib
// (reference src/kernels/factor.hxx:149-263, src/kernels/factor_indef.hxx:23-49,
// 201-342; GPU variants src/StarPU/cuda/kernels.hxx:45-182) and, with the
// inverted diagonal block as B, solve_block (src/kernels/factor.hxx:95-138).
//
// Both operands are column-major panels (row index contiguous), staged into
// shared memory k-column by k-column with TMA bulk copies through a 4-stage
// mbarrier pipeline by one producer warp; 8 consumer warps (4 x 2, 32 x 64 per
// warp) issue DMMA.8x8x4 from conflict-free padded shared memory.
// ---------------------------------------------------------------------------
constexpr int GT_BM = 128;             // tile rows
constexpr int GT_BN = 128;             // tile cols
constexpr int GT_KT = 16;              // k-columns per stage
constexpr int GT_LDS = 132;            // padded smem row stride (132 mod 16 == 4: conflict free)
constexpr int GT_STAGES = 4;
constexpr int GT_CONSUMERS = 16;       // warps, 4 (M) x 4 (N), 32 x 32 each: 4 per SM sub-partition
constexpr int GT_THREADS = (GT_CONSUMERS + 1) * 32;
constexpr int GT_LDC = 130;            // epilogue staging stride (130 mod 8 == 2: conflict free)
constexpr size_t GT_STAGE_DOUBLES = 2 * GT_KT * GT_LDS;
constexpr size_t GT_SMEM_BYTES = GT_STAGES * GT_STAGE_DOUBLES * sizeof(double) + 2 * GT_STAGES * sizeof(uint64_t);
static_assert((size_t)GT_BN * GT_LDC <= GT_STAGES * GT_STAGE_DOUBLES, "epilogue tile must fit in the pipeline buffers");

struct GemmTile {
   const double* A;   // first row of the A operand tile, k = 0 column
   const double* B;   // first row of the B operand tile, k = 0 column
   int lda, ldb;      // column strides (doubles)
   int arows, brows;  // valid rows (<=128) in each operand tile (rounded up to even inside)
   int K;             // depth
   // destination tile (for the L2 prefetch issued by the producer warp): column c of the
   // tile starts at dst + c*ldd; rows [0,128) -- only when prefetch is set
   const double* dst;
   int ldd;
   bool prefetch;
   int pf_c0, pf_c1;  // tile-relative column range worth prefetching
   int pf_lines;      // 128 B lines per column
   // fused extend-add (contribution tiles): the children's entries the epilogue gathers are
   // pulled into L2 by the producer warp while the tensor cores work
   const double* gsrc[2];
   const int* gmap[2];    // parent contribution row -> child row (-1: none), increasing
   int gld[2];
   int g_r0, g_c0;        // contribution-relative row / column of the tile origin
   int g_k;               // order of the parent's contribution block
};

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(GT_CONSUMERS * 32) : "memory"); }

// Runs the pipeline; on return (consumer warps) the 128 x 128 product tile is staged in
// shared memory as Cs[c*GT_LDC + r] and the consumer warps are synchronised.  Returns false
// for the producer warp (which has nothing left to do).
__device__ __forceinline__ bool gemm_tile_mainloop(const GemmTile& t, double* smem) {
   uint64_t* full = reinterpret_cast<uint64_t*>(smem + GT_STAGES * GT_STAGE_DOUBLES);
   uint64_t* empty = full + GT_STAGES;
   const int warp = threadIdx.x >> 5;
   const int lane = threadIdx.x & 31;
   if (threadIdx.x == 0) {
      for (int s = 0; s < GT_STAGES; ++s) {
         mbar_init(&full[s], 1);
         mbar_init(&empty[s], GT_CONSUMERS);
      }
      mbar_fence_init();
   }
   __syncthreads();
   const int nk = (t.K + GT_KT - 1) / GT_KT;
   const bool same = (t.A == t.B) && (t.lda == t.ldb);   // diagonal tile: stage the panel once
   if (warp == GT_CONSUMERS) {
      // ===== producer warp: lane l < 16 copies A column l, lane 16+l copies B column l =====
      const uint32_t abytes = (uint32_t)(((t.arows + 1) & ~1) * sizeof(double));
      const uint32_t bbytes = (uint32_t)(((t.brows + 1) & ~1) * sizeof(double));
      for (int kb = 0; kb < nk; ++kb) {
         const int s = kb % GT_STAGES;
         const uint32_t ph = (kb / GT_STAGES) & 1;
         mbar_wait(&empty[s], ph ^ 1);
         const int k0 = kb * GT_KT;
         const int kv = min(GT_KT, t.K - k0);           // valid k-columns in this stage
         double* As = smem + s * GT_STAGE_DOUBLES;
         double* Bs = As + GT_KT * GT_LDS;
         const int kc = lane & 15;
         const bool isB = lane >= 16;
         if (kc >= kv && kc < ((kv + 3) & ~3)) {
            // zero-fill the tail of a partial k-group (generic proxy; ordered by the arrive below)
            double* dst = (isB ? Bs : As) + kc * GT_LDS;
            if (!(isB && same))
               for (int r = 0; r < GT_BM; ++r) dst[r] = 0.0;
         }
         __syncwarp();
         if (lane == 0) mbar_expect_tx(&full[s], kv * (abytes + (same ? 0u : bbytes)));
         __syncwarp();
         if (kc < kv) {
            if (!isB)
               tma_bulk_g2s(As + kc * GT_LDS, t.A + (size_t)(k0 + kc) * t.lda, abytes, &full[s]);
            else if (!same)
               tma_bulk_g2s(Bs + kc * GT_LDS, t.B + (size_t)(k0 + kc) * t.ldb, bbytes, &full[s]);
         }
         if (kb == min(nk, GT_STAGES) - 1 && (t.gsrc[0] || t.gsrc[1])) {
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
               if (!t.gsrc[s2]) continue;
               // child rows hit by this tile's rows: [rlo, rhi]
               int rlo = 0x7fffffff, rhi = -1;
#pragma unroll
               for (int q = 0; q < 4; ++q) {
                  const int r = t.g_r0 + lane + 32 * q;
                  const int v = (r >= 0 && r < t.g_k) ? t.gmap[s2][r] : -1;
                  if (v >= 0) { rlo = min(rlo, v); rhi = max(rhi, v); }
               }
#pragma unroll
               for (int o = 16; o > 0; o >>= 1) {
                  rlo = min(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
                  rhi = max(rhi, __shfl_xor_sync(0xffffffffu, rhi, o));
               }
               if (rhi < 0) continue;
               for (int cc = lane; cc < GT_BN; cc += 32) {
                  const int c = t.g_c0 + cc;
                  const int ic = (c >= 0 && c < t.g_k) ? t.gmap[s2][c] : -1;
                  if (ic < 0) continue;
                  const int lo = max(rlo, ic);
                  if (lo > rhi) continue;
                  const char* p0 = reinterpret_cast<const char*>(t.gsrc[s2] + (size_t)ic * t.gld[s2] + lo);
                  const char* p1 = reinterpret_cast<const char*>(t.gsrc[s2] + (size_t)ic * t.gld[s2] + rhi);
                  for (const char* p = p0; p <= p1; p += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(p1));
               }
            }
         }
         if (kb == min(nk, GT_STAGES) - 1 && t.prefetch) {
            // pull the destination tile into L2 while the tensor cores work
            for (int c = t.pf_c0 + lane; c < t.pf_c1; c += 32) {
               const char* p = reinterpret_cast<const char*>(t.dst + (size_t)c * t.ldd);
               for (int q = 0; q < t.pf_lines; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + q * 128));
            }
         }
      }
      return false;
   }
   // ===== consumer warps =====
   const int wm = warp & 3, wn = warp >> 2;
   const int g = lane >> 2, tq = lane & 3;
   double acc[4][4][2];
#pragma unroll
   for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
   for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % GT_STAGES;
      const uint32_t ph = (kb / GT_STAGES) & 1;
      mbar_wait(&full[s], ph);
      const double* As = smem + s * GT_STAGE_DOUBLES;
      const double* Bs = same ? As : As + GT_KT * GT_LDS;
      const int kv = min(GT_KT, t.K - kb * GT_KT);
      const int kg = (kv + 3) >> 2;
      const double* ap = As + tq * GT_LDS + wm * 32 + g;
      const double* bp = Bs + tq * GT_LDS + wn * 32 + g;
#pragma unroll 1
      for (int k4 = 0; k4 < kg; ++k4) {
         double a[4], b[4];
#pragma unroll
         for (int i = 0; i < 4; ++i) a[i] = ap[i * 8];
#pragma unroll
         for (int j = 0; j < 4; ++j) b[j] = bp[j * 8];
#pragma unroll
         for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
         ap += 4 * GT_LDS;
         bp += 4 * GT_LDS;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
   }
   // every stage has been consumed by this warp; once all consumers are here the pipeline
   // buffers are dead and can hold the output tile
   consumer_bar();
#pragma unroll
   for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
         double* col = smem + (size_t)(wn * 32 + j * 8 + 2 * tq + h) * GT_LDC + wm * 32 + g;
#pragma unroll
         for (int i = 0; i < 4; ++i) col[i * 8] = acc[i][j][h];
      }
   consumer_bar();
   return true;
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

// mode 0: trailing update inside the L panel after block column [p0, p0+pw):
//         L[r][c] -= sum_k L[r][k] L[c][k],  c in [p0+pw, n), r in [c, m)
// mode 1: contribution block:  C[r-n][c-n] = beta*C - sum_{k<n} L[r][k] L[c][k], n <= c <= r < m
// mode 2: panel solve with the inverted diagonal block W (pw x pw, ld wld):
//         L[r][p0+c] = sum_k L[r][p0+k] W[c][k],  r in [p0+pw, m)
// nb is the block-column width; `step` the block column index.  Modes 0 and 1 enumerate the
// tile columns tstart, tstart + tstep, ... of the tile grid (look-ahead scheduling: the grid
// size limits a launch to the first tile column, tstart = 1 skips it; split fronts: the tile
// columns this rank owns).
static __global__ void __launch_bounds__(GT_THREADS, 1)
k_gemm_batched(DevTree T, TileBatch batch, int mode, int step, int nb, const double* __restrict__ W, int wld,
               int tstart, int tstep) {
   extern __shared__ __align__(128) double smem[];
   const int item = blockIdx.x;
   const int fi = find_front(batch, item);
   const int f = batch.fronts[fi];
   int local = item - batch.prefix[fi];
   const int m = T.m[f], n = T.n[f], ldl = T.ldl[f];
   double* Lf = T.L + T.loff[f];
   const int p0 = step * nb;
   const int pw = min(nb, n - p0);

   GemmTile t;
   int i0, j0;          // absolute front row / column of the tile origin
   // destination addressing: element (r, c) of the tile (absolute front indices) lives at
   // dbase + c*ldd + r ; valid iff  r in [rlo(c), m)  and  c in [clo, chi)
   double* dbase;
   int ldd, clo, chi, rmin;
   bool lower;          // additionally require r >= c
   int op;              // 0: dst -= v   1: dst = -v   2: dst = v
   if (mode >= 3) {
      // ---- indefinite path: operands are W = L*D (A side) and L (B side); the extent of the
      // update is read from the device-resident front state (nothing here is known to the host)
      // mode 3: block column [kbeg, kbeg+klen) -> the rest of the current outer panel
      // mode 5: outer panel [obeg, p0) -> everything behind the panel
      // mode 4: all eliminated columns -> contribution block
      const FrontState st = T.state[f];
      const double* Wf = T.W + T.woff[f];
      const int kbeg = (mode == 3) ? st.kbeg : (mode == 5 ? st.obeg : 0);
      t.K = (mode == 3) ? st.klen : (mode == 5 ? st.p0 - st.obeg : st.nelim);
      if (mode != 4 && t.K == 0) return;
      const int first = (mode == 3) ? st.p0 : (mode == 5 ? st.oend : n);     // first column/row of the updated region
      const int base = first & ~1;
      const int cend = (mode == 3) ? st.oend : (mode == 5 ? n : m);
      if (cend <= first) return;
      const int TR = (m - base + GT_BM - 1) / GT_BM;
      const int TC = (cend - base + GT_BN - 1) / GT_BN;
      int tj = 0;
      while (tj < TC && local >= TR - tj) { local -= TR - tj; ++tj; }
      if (tj >= TC) return;
      const int ti = tj + local;
      i0 = base + ti * GT_BM;
      j0 = base + tj * GT_BN;
      t.A = Wf + (size_t)kbeg * ldl + i0;
      t.B = Lf + (size_t)kbeg * ldl + j0;
      t.lda = t.ldb = ldl;
      t.arows = min(GT_BM, m - i0);
      t.brows = min(GT_BN, m - j0);
      if (mode != 4) {
         dbase = Lf; ldd = ldl; clo = first; chi = cend; rmin = 0; lower = true; op = 0;
         t.prefetch = true;
      } else {
         const int ldc = T.ldc[f];
         dbase = T.C + T.coff[f] - (size_t)n * ldc - n; ldd = ldc; clo = n; chi = m; rmin = 0; lower = true;
         op = 1;      // children are extend-added afterwards (k_assemble_indef part 1)
         t.prefetch = false;
      }
   } else if (mode == 2) {
      // tile origin rounded down to an even row so the TMA source stays 16 B aligned
      i0 = ((p0 + pw) & ~1) + local * GT_BM;
      j0 = p0;
      t.A = Lf + (size_t)p0 * ldl + i0;
      t.lda = ldl;
      t.arows = min(GT_BM, m - i0);
      t.B = W + (size_t)fi * wld * wld;   // slot fi of this launch's inverse buffer
      t.ldb = wld;
      t.brows = pw;
      t.K = pw;
      dbase = Lf; ldd = ldl; clo = p0; chi = p0 + pw; rmin = p0 + pw; lower = false; op = 2;
      t.prefetch = false;
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
      j0 = base + tj * GT_BN;
      const int kbeg = (mode == 0) ? p0 : 0;
      t.K = (mode == 0) ? pw : n;
      t.A = Lf + (size_t)kbeg * ldl + i0;
      t.B = Lf + (size_t)kbeg * ldl + j0;
      t.lda = t.ldb = ldl;
      t.arows = min(GT_BM, m - i0);
      t.brows = min(GT_BN, m - j0);
      if (mode == 0) {
         dbase = Lf; ldd = ldl; clo = p0 + pw; chi = n; rmin = 0; lower = true; op = 0;
         t.prefetch = true;
      } else {
         const int ldc = T.ldc[f];
         dbase = T.C + T.coff[f] - (size_t)n * ldc - n; ldd = ldc; clo = n; chi = m; rmin = 0; lower = true;
         op = 1;      // children are extend-added afterwards (k_assemble part 1)
         t.prefetch = false;
      }
   }
   t.dst = dbase + (size_t)j0 * ldd + i0;
   t.ldd = ldd;
   t.pf_c0 = max(0, clo - j0);
   t.pf_c1 = min(GT_BN, chi - j0);
   t.pf_lines = (min(GT_BM, m - i0) * 8 + 127) / 128;
   t.gsrc[0] = t.gsrc[1] = nullptr;
   if (op == 1 && T.fchild) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
         const int fc = T.fchild[2 * f + s];
         if (fc < 0) continue;
         t.gsrc[s] = T.C + T.coff[fc];
         t.gld[s] = T.ldc[fc];
         t.gmap[s] = T.pinv + T.pinvoff[2 * f + s];
      }
      t.g_r0 = i0 - n; t.g_c0 = j0 - n; t.g_k = m - n;
   }
   if (!gemm_tile_mainloop(t, smem)) return;

   // ---- coalesced epilogue: warp w owns tile columns w*8 .. w*8+7, lanes stride the rows ----
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const bool interior = (i0 + GT_BM <= m) && (j0 >= clo) && (j0 + GT_BN <= chi) && (i0 >= j0 + GT_BN - 1 || !lower) &&
                         (i0 >= rmin);
   // Contribution block: fused extend-add.  The children's generated elements that the
   // reference adds afterwards (assemble_contrib_block, src/kernels/assemble.hxx:343-517)
   // are gathered here through the parent-row -> child-row maps, so the block is written
   // once instead of written, re-read and re-written (8 + 8 B per entry instead of 8 + 24).
   // All map entries a warp needs (4 row groups, its 8 columns) are loaded up front in one
   // batch, and the gathers are predicated loads without branches, so that a warp keeps 16
   // of them in flight instead of paying one memory latency per column.
   const bool fused = t.gsrc[0] || t.gsrc[1];
   int fir[2][4], fic[2][8];
   if (fused) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
#pragma unroll
         for (int q = 0; q < 4; ++q) {
            const int r = i0 + lane + 32 * q;
            fir[s][q] = (t.gsrc[s] && r >= n && r < m) ? t.gmap[s][r - n] : -1;
         }
#pragma unroll
         for (int u = 0; u < 8; ++u) {
            const int c = j0 + warp * 8 + u;
            fic[s][u] = (t.gsrc[s] && c >= n && c < m) ? t.gmap[s][c - n] : -1;
         }
      }
   }
#pragma unroll
   for (int cq = 0; cq < 8; cq += 4) {
      double v[4][4], d[4][4];
      double* gp[4];
      bool cok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
         const int ct = warp * 8 + cq + u;          // column inside the tile
         const int c = j0 + ct;
         cok[u] = interior || (c >= clo && c < chi);
         gp[u] = dbase + (size_t)c * ldd + i0;
#pragma unroll
         for (int q = 0; q < 4; ++q) v[u][q] = smem[(size_t)ct * GT_LDC + lane + 32 * q];
      }
      if (fused) {
#pragma unroll
         for (int s = 0; s < 2; ++s) {
            if (!t.gsrc[s]) continue;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
               const int ic = fic[s][cq + u];
               const double* col = t.gsrc[s] + (size_t)max(ic, 0) * t.gld[s];
#pragma unroll
               for (int q = 0; q < 4; ++q) {
                  const bool hit = ic >= 0 && fir[s][q] >= ic;      // maps are increasing: row >= col
                  const double gv = hit ? col[fir[s][q]] : 0.0;
                  v[u][q] -= gv;
               }
            }
         }
      }
      if (op == 0) {
#pragma unroll
         for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
               const int r = i0 + lane + 32 * q;
               const bool ok = interior || (cok[u] && r < m && r >= rmin && (!lower || r >= j0 + warp * 8 + cq + u));
               d[u][q] = ok ? gp[u][lane + 32 * q] : 0.0;
            }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
         for (int q = 0; q < 4; ++q) {
            const int r = i0 + lane + 32 * q;
            const bool ok = interior || (cok[u] && r < m && r >= rmin && (!lower || r >= j0 + warp * 8 + cq + u));
            if (ok) gp[u][lane + 32 * q] = (op == 0) ? d[u][q] - v[u][q] : (op == 1 ? -v[u][q] : v[u][q]);
         }
   }
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

}  // namespace sylver_b200
