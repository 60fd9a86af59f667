// Triangular-solve kernels: level-batched, one CTA per front, deterministic
// (children's update vectors are added by the parent in a fixed order instead
// of atomics).  Replaces NumericTree::solve_fwd / solve_diag_bwd and the
// per-front ldlt_app_solve_* / cholesky_solve_* (reference
// src/NumericTree.hxx:408-532, spral/src/ssids/cpu/kernels/ldlt_app.cxx:2536-2590).
#pragma once
#include "kernels.cuh"

namespace sylver_b200 {

constexpr int SOLVE_THREADS = 512;

struct SolveArgs {
   DevTree T;
   const int* rlist;        // 1-based global row indices, concatenated
   const long* rptr;        // 1-based offsets into rlist
   const int* perm;         // indefinite: per-front pivot permutation (concatenated, permoff), 1-based
   const long* permoff;
   const double* D;         // indefinite: D^-1 storage, 2 per column, at doff[f]
   const long* doff;
   const int* nelim;        // indefinite: eliminated columns per front
   const long* xwoff;       // per-front offset into xw
   double* xw;              // work vectors (sum of m)
   const int* child_ptr;
   const int* child_list;
   double* x;               // right-hand side / solution in elimination order
   int posdef;
};

// index of local variable i (< n) of front f in the global vector
__device__ __forceinline__ int var_index(const SolveArgs& a, int f, int i, const int* rl) {
   if (a.posdef) return rl[i] - 1;
   return a.perm[a.permoff[f] + i] - 1;
}

static __global__ void __launch_bounds__(SOLVE_THREADS) k_solve_fwd(SolveArgs a, const int* __restrict__ fronts) {
   const int f = fronts[blockIdx.x];
   const int m = a.T.m[f], n = a.T.n[f], ldl = a.T.ldl[f];
   const int ne = a.posdef ? n : a.nelim[f];
   const double* L = a.T.L + a.T.loff[f];
   double* xw = a.xw + a.xwoff[f];
   const int* rl = a.rlist + (a.rptr[f] - 1);
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   // Children's updates arrive in the front's ORIGINAL row order (cmap); pivoting permuted the
   // fully-summed rows, so they are first summed in original order, folded into the global
   // vector at the front's own variables (rl[j], exclusively ours), and only then gathered in
   // pivot order.  Rows >= n are untouched by pivoting (shifted by the delays).
   const int ncol0 = (int)(a.rptr[f + 1] - a.rptr[f]) - (m - n);
   const int nd = n - ncol0;
   for (int i = tid; i < m; i += SOLVE_THREADS) xw[i] = 0.0;
   __syncthreads();
   for (int ci = a.child_ptr[f]; ci < a.child_ptr[f + 1]; ++ci) {
      const int c = a.child_list[ci];
      const int cn = a.T.n[c];
      const int k = a.T.m[c] - cn;
      const int* cm = a.T.cmap + a.T.cmapoff[c];
      const double* src = a.xw + a.xwoff[c];
      for (int i = tid; i < k; i += SOLVE_THREADS) {
         const int r = cm[i];
         xw[r < ncol0 ? r : r + nd] += src[cn + i];
      }
      __syncthreads();
   }
   for (int j = tid; j < ncol0; j += SOLVE_THREADS) a.x[rl[j] - 1] += xw[j];
   __syncthreads();
   for (int i = tid; i < n; i += SOLVE_THREADS) xw[i] = a.x[var_index(a, f, i, rl)];
   __syncthreads();
   for (int j0 = 0; j0 < ne; j0 += 32) {
      const int jb = min(32, ne - j0);
      if (warp == 0) {
         double y = (lane < jb) ? xw[j0 + lane] : 0.0;
         for (int k = 0; k < jb; ++k) {
            double yk = __shfl_sync(0xffffffffu, y, k);
            if (a.posdef) yk /= L[(size_t)(j0 + k) * ldl + j0 + k];
            if (lane == k) y = yk;
            if (lane > k && lane < jb) y -= L[(size_t)(j0 + k) * ldl + j0 + lane] * yk;
         }
         if (lane < jb) xw[j0 + lane] = y;
      }
      __syncthreads();
      for (int r = j0 + jb + tid; r < m; r += SOLVE_THREADS) {
         double s = 0.0;
         const double* lp = L + (size_t)j0 * ldl + r;
         for (int k = 0; k < jb; ++k) s += lp[(size_t)k * ldl] * xw[j0 + k];
         xw[r] -= s;
      }
      __syncthreads();
   }
   for (int i = tid; i < n; i += SOLVE_THREADS) a.x[var_index(a, f, i, rl)] = xw[i];
}

static __global__ void __launch_bounds__(SOLVE_THREADS) k_solve_bwd(SolveArgs a, const int* __restrict__ fronts) {
   const int f = fronts[blockIdx.x];
   const int m = a.T.m[f], n = a.T.n[f], ldl = a.T.ldl[f];
   const int ne = a.posdef ? n : a.nelim[f];
   const double* L = a.T.L + a.T.loff[f];
   double* xw = a.xw + a.xwoff[f];
   const int* rl = a.rlist + (a.rptr[f] - 1);
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   constexpr int NW = SOLVE_THREADS / 32;
   const int ncol0 = (int)(a.rptr[f + 1] - a.rptr[f]) - (m - n);   // original fully-summed columns
   for (int i = tid; i < m; i += SOLVE_THREADS)
      xw[i] = (i < n) ? a.x[var_index(a, f, i, rl)] : a.x[rl[ncol0 + (i - n)] - 1];
   __syncthreads();
   const int nblk = (ne + 31) / 32;
   for (int b = nblk - 1; b >= 0; --b) {
      const int j0 = b * 32;
      const int jb = min(32, ne - j0);
      for (int kk = warp; kk < jb; kk += NW) {
         const double* lp = L + (size_t)(j0 + kk) * ldl;
         double s = 0.0;
         for (int r = j0 + jb + lane; r < m; r += 32) s += lp[r] * xw[r];
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
         if (lane == 0) xw[j0 + kk] -= s;
      }
      __syncthreads();
      if (warp == 0) {
         double y = (lane < jb) ? xw[j0 + lane] : 0.0;
         for (int k = jb - 1; k >= 0; --k) {
            double yk = __shfl_sync(0xffffffffu, y, k);
            if (a.posdef) yk /= L[(size_t)(j0 + k) * ldl + j0 + k];
            if (lane == k) y = yk;
            if (lane < k) y -= L[(size_t)(j0 + lane) * ldl + j0 + k] * yk;
         }
         if (lane < jb) xw[j0 + lane] = y;
      }
      __syncthreads();
   }
   for (int i = tid; i < ne; i += SOLVE_THREADS) a.x[var_index(a, f, i, rl)] = xw[i];
}

// ---------------------------------------------------------------------------
// The same solves with G CTAs per front, for the levels that hold few, large fronts (the root of
// lap27_100 alone is 0.9 GB of L that one CTA would stream at the bandwidth of one SM).  The
// columns are processed in blocks of 128: CTA 0 of the group solves the triangular block (in
// shared memory), the G CTAs share the rows below it (forward: one thread per row; backward:
// partial dot products over contiguous row ranges, summed by CTA 0 in a fixed order), with a
// group barrier after each half (counter `bar[front slot]`, zeroed by the host before the launch;
// the G CTAs of a front have consecutive block indices).  Values that cross CTAs are read with
// L1-bypassing loads.
// ---------------------------------------------------------------------------
constexpr int SOLVE_BW = 128;
constexpr int SOLVE_GMAX = 64;

struct SolveGroup {
   int* bar;
   int g, G, epoch;
};
__device__ __forceinline__ void solve_barrier(SolveGroup& sg) {
   __syncthreads();
   ++sg.epoch;
   if (sg.G > 1 && threadIdx.x == 0) {
      __threadfence();
      atomicAdd(sg.bar, 1);
      while (atomicAdd(sg.bar, 0) < sg.G * sg.epoch) __nanosleep(32);
      __threadfence();
   }
   __syncthreads();
}

static __global__ void __launch_bounds__(SOLVE_THREADS) k_solve_fwd_multi(SolveArgs a, const int* __restrict__ fronts,
                                                                          int G, int* __restrict__ bar) {
   __shared__ double sx[SOLVE_BW];
   const int fi = blockIdx.x / G;
   SolveGroup sg{bar + fi, (int)(blockIdx.x % G), G, 0};
   const int f = fronts[fi];
   const int m = a.T.m[f], n = a.T.n[f], ldl = a.T.ldl[f];
   const int ne = a.posdef ? n : a.nelim[f];
   const double* L = a.T.L + a.T.loff[f];
   double* xw = a.xw + a.xwoff[f];
   const int* rl = a.rlist + (a.rptr[f] - 1);
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   if (sg.g == 0) {
      // gather: children's updates in original row order, own variables in pivot order (k_solve_fwd)
      const int ncol0 = (int)(a.rptr[f + 1] - a.rptr[f]) - (m - n);
      const int nd = n - ncol0;
      for (int i = tid; i < m; i += SOLVE_THREADS) xw[i] = 0.0;
      __syncthreads();
      for (int ci = a.child_ptr[f]; ci < a.child_ptr[f + 1]; ++ci) {
         const int c = a.child_list[ci];
         const int cn = a.T.n[c];
         const int k = a.T.m[c] - cn;
         const int* cm = a.T.cmap + a.T.cmapoff[c];
         const double* src = a.xw + a.xwoff[c];
         for (int i = tid; i < k; i += SOLVE_THREADS) {
            const int r = cm[i];
            xw[r < ncol0 ? r : r + nd] += __ldcg(src + cn + i);
         }
         __syncthreads();
      }
      for (int j = tid; j < ncol0; j += SOLVE_THREADS) a.x[rl[j] - 1] += xw[j];
      __syncthreads();
      for (int i = tid; i < n; i += SOLVE_THREADS) xw[i] = a.x[var_index(a, f, i, rl)];
   }
   solve_barrier(sg);
   const int t0 = sg.g * SOLVE_THREADS + tid, nt = G * SOLVE_THREADS;
   for (int j0 = 0; j0 < ne; j0 += SOLVE_BW) {
      const int jb = min(SOLVE_BW, ne - j0);
      if (sg.g == 0) {
         // ---- triangular block in shared memory, 32 columns at a time ----
         if (tid < jb) sx[tid] = __ldcg(xw + j0 + tid);
         __syncthreads();
         for (int q0 = 0; q0 < jb; q0 += 32) {
            const int qb = min(32, jb - q0);
            if (warp == 0) {
               double y = (lane < qb) ? sx[q0 + lane] : 0.0;
               for (int k = 0; k < qb; ++k) {
                  double yk = __shfl_sync(0xffffffffu, y, k);
                  if (a.posdef) yk /= L[(size_t)(j0 + q0 + k) * ldl + j0 + q0 + k];
                  if (lane == k) y = yk;
                  if (lane > k && lane < qb) y -= L[(size_t)(j0 + q0 + k) * ldl + j0 + q0 + lane] * yk;
               }
               if (lane < qb) sx[q0 + lane] = y;
            }
            __syncthreads();
            // rows of this block below the 32 columns: 4 threads per row
            {
               const int r = q0 + qb + (tid >> 2), part = tid & 3;
               double s = 0.0;
               if (r < jb) {
                  const double* lp = L + (size_t)(j0 + q0) * ldl + j0 + r;
                  for (int k = part; k < qb; k += 4) s += lp[(size_t)k * ldl] * sx[q0 + k];
               }
               s += __shfl_xor_sync(0xffffffffu, s, 1);
               s += __shfl_xor_sync(0xffffffffu, s, 2);
               if (r < jb && part == 0) sx[r] -= s;
            }
            __syncthreads();
         }
         if (tid < jb) xw[j0 + tid] = sx[tid];
      }
      solve_barrier(sg);
      // ---- rows below the block, shared by the group: one thread per row ----
      if (sg.g != 0) {
         if (tid < jb) sx[tid] = __ldcg(xw + j0 + tid);
         __syncthreads();
      }
      for (int r = j0 + jb + t0; r < m; r += nt) {
         double s = 0.0;
         const double* lp = L + (size_t)j0 * ldl + r;
#pragma unroll 4
         for (int k = 0; k < jb; ++k) s += lp[(size_t)k * ldl] * sx[k];
         xw[r] = __ldcg(xw + r) - s;
      }
      solve_barrier(sg);
   }
   if (sg.g == 0)
      for (int i = tid; i < n; i += SOLVE_THREADS) a.x[var_index(a, f, i, rl)] = __ldcg(xw + i);
}

// part: [front slot][G][SOLVE_BW] partial dot products
static __global__ void __launch_bounds__(SOLVE_THREADS) k_solve_bwd_multi(SolveArgs a, const int* __restrict__ fronts,
                                                                          int G, int* __restrict__ bar,
                                                                          double* __restrict__ part) {
   __shared__ double sx[SOLVE_BW];
   const int fi = blockIdx.x / G;
   SolveGroup sg{bar + fi, (int)(blockIdx.x % G), G, 0};
   const int f = fronts[fi];
   const int m = a.T.m[f], n = a.T.n[f], ldl = a.T.ldl[f];
   const int ne = a.posdef ? n : a.nelim[f];
   const double* L = a.T.L + a.T.loff[f];
   double* xw = a.xw + a.xwoff[f];
   const int* rl = a.rlist + (a.rptr[f] - 1);
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   constexpr int NW = SOLVE_THREADS / 32;
   double* mypart = part + ((size_t)fi * G + sg.g) * SOLVE_BW;
   const double* fpart = part + (size_t)fi * G * SOLVE_BW;
   if (sg.g == 0) {
      const int ncol0 = (int)(a.rptr[f + 1] - a.rptr[f]) - (m - n);   // original fully-summed columns
      for (int i = tid; i < m; i += SOLVE_THREADS)
         xw[i] = (i < n) ? a.x[var_index(a, f, i, rl)] : a.x[rl[ncol0 + (i - n)] - 1];
   }
   solve_barrier(sg);
   const int nblk = (ne + SOLVE_BW - 1) / SOLVE_BW;
   for (int b = nblk - 1; b >= 0; --b) {
      const int j0 = b * SOLVE_BW;
      const int jb = min(SOLVE_BW, ne - j0);
      // ---- partial dot products over this CTA's contiguous share of the rows below the block ----
      {
         const int rows = m - (j0 + jb);
         const int per = (rows + G - 1) / G;
         const int lo = j0 + jb + sg.g * per, hi = min(m, lo + per);
         for (int kk = warp; kk < jb; kk += NW) {
            const double* lp = L + (size_t)(j0 + kk) * ldl;
            double s = 0.0;
            for (int r = lo + lane; r < hi; r += 32) s += lp[r] * __ldcg(xw + r);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) mypart[kk] = s;
         }
      }
      solve_barrier(sg);
      if (sg.g == 0) {
         if (tid < jb) {
            double v = __ldcg(xw + j0 + tid);
            for (int q = 0; q < G; ++q) v -= __ldcg(fpart + (size_t)q * SOLVE_BW + tid);      // fixed order
            sx[tid] = v;
         }
         __syncthreads();
         // ---- triangular block, 32 columns at a time from the last ----
         const int nq = (jb + 31) / 32;
         for (int qi = nq - 1; qi >= 0; --qi) {
            const int q0 = qi * 32;
            const int qb = min(32, jb - q0);
            // rows of this block below the 32 columns (already solved): warp per column
            for (int kk = warp; kk < qb; kk += NW) {
               const double* lp = L + (size_t)(j0 + q0 + kk) * ldl + j0;
               double s = 0.0;
               for (int r = q0 + qb + lane; r < jb; r += 32) s += lp[r] * sx[r];
#pragma unroll
               for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
               if (lane == 0) sx[q0 + kk] -= s;
            }
            __syncthreads();
            if (warp == 0) {
               double y = (lane < qb) ? sx[q0 + lane] : 0.0;
               for (int k = qb - 1; k >= 0; --k) {
                  double yk = __shfl_sync(0xffffffffu, y, k);
                  if (a.posdef) yk /= L[(size_t)(j0 + q0 + k) * ldl + j0 + q0 + k];
                  if (lane == k) y = yk;
                  if (lane < k) y -= L[(size_t)(j0 + q0 + lane) * ldl + j0 + q0 + k] * yk;
               }
               if (lane < qb) sx[q0 + lane] = y;
            }
            __syncthreads();
         }
         if (tid < jb) xw[j0 + tid] = sx[tid];
      }
      solve_barrier(sg);
   }
   if (sg.g == 0)
      for (int i = tid; i < ne; i += SOLVE_THREADS) a.x[var_index(a, f, i, rl)] = __ldcg(xw + i);
}

// ---- multi-GPU solve (positive definite): x is replicated; the variables a front eliminates
// are the contiguous range x[v0, v0 + ncol), v0 = rlist[rptr[f]-1] - 1 ----
// pack: buf[off[i] + j] = (front owned by `me`) ? x[v0 + j] : 0, one CTA per front of the level;
// after an all-reduce(sum) of buf every rank unpacks the freshly solved entries.
static __global__ void __launch_bounds__(256) k_pack_level(const int* __restrict__ fronts, const int* __restrict__ off,
                                                           const int* __restrict__ owner, int me,
                                                           const int* __restrict__ rlist, const long* __restrict__ rptr,
                                                           const int* __restrict__ ncol, const double* __restrict__ x,
                                                           double* __restrict__ buf, int unpack, double* __restrict__ xout) {
   const int f = fronts[blockIdx.x];
   const int v0 = rlist[rptr[f] - 1] - 1;
   const int nc = ncol[f];
   double* b = buf + off[blockIdx.x];
   if (unpack) {
      for (int j = threadIdx.x; j < nc; j += blockDim.x) xout[v0 + j] = b[j];
   } else {
      const bool mine = owner[f] == me;
      for (int j = threadIdx.x; j < nc; j += blockDim.x) b[j] = mine ? x[v0 + j] : 0.0;
   }
}
static __global__ void __launch_bounds__(256) k_zero_unowned(int nfronts, const int* __restrict__ owner, int me,
                                                             const int* __restrict__ rlist, const long* __restrict__ rptr,
                                                             const int* __restrict__ ncol, double* __restrict__ x) {
   const int f = blockIdx.x;
   if (f >= nfronts || owner[f] == me) return;
   const int v0 = rlist[rptr[f] - 1] - 1;
   for (int j = threadIdx.x; j < ncol[f]; j += blockDim.x) x[v0 + j] = 0.0;
}

// D^-1 application of every front (indefinite only; ldlt_app_solve_diag,
// spral/src/ssids/cpu/kernels/ldlt_app.cxx:2556-2578).  One warp per front.
static __global__ void __launch_bounds__(256) k_solve_diag(SolveArgs a, int nfronts, const int* __restrict__ owner,
                                                           int me) {
   const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   if (f >= nfronts) return;
   if (owner && owner[f] != me) return;      // multi-rank: D^-1 lives with the front's owner
   const int lane = threadIdx.x & 31;
   const int ne = a.nelim[f];
   const double* d = a.D + a.doff[f];
   const int* perm = a.perm + a.permoff[f];
   double* xw = a.xw + a.xwoff[f];
   for (int i = lane; i < ne; i += 32) {
      const bool second = (i > 0) && isinf(d[2 * i]);
      const bool first = (i + 1 < ne) && isinf(d[2 * i + 2]);
      const double xi = a.x[perm[i] - 1];
      double r;
      if (second) r = d[2 * i - 1] * a.x[perm[i - 1] - 1] + d[2 * i + 1] * xi;
      else if (first) r = d[2 * i] * xi + d[2 * i + 1] * a.x[perm[i + 1] - 1];
      else r = d[2 * i] * xi;
      xw[i] = r;          // staged: 2x2 partners may be handled by another pass of the loop
   }
   __syncwarp();
   for (int i = lane; i < ne; i += 32) a.x[perm[i] - 1] = xw[i];
}

// ---- multi-rank solve with delayed pivots: x is replicated, but the variables a front
// touches are no longer its own contiguous range (delayed columns travel up the tree with
// their variable), so after every level each rank publishes the entries it changed:
// buf[i] = changed ? x[i] : 0, buf[n+i] = changed ? 1 : 0; all-reduce(sum); whoever changed an
// entry (at most one rank: fronts of a level touch disjoint variables) wins.
static __global__ void k_delta_pack(int n, const double* __restrict__ x, const double* __restrict__ xprev,
                                    double* __restrict__ buf) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   const bool ch = __double_as_longlong(x[i]) != __double_as_longlong(xprev[i]);
   buf[i] = ch ? x[i] : 0.0;
   buf[n + i] = ch ? 1.0 : 0.0;
}
static __global__ void k_delta_unpack(int n, double* __restrict__ x, double* __restrict__ xprev,
                                      const double* __restrict__ buf) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   if (buf[n + i] != 0.0) x[i] = buf[i];
   xprev[i] = x[i];
}

}  // namespace sylver_b200
