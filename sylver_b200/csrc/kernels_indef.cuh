// Device kernels of the indefinite (APTP) path: a-posteriori threshold pivoting
// LDL^T with 1x1/2x2 pivots and delayed columns (sm_100a).
//
// Reference algorithm: FactorIndefAPP::factor_front_indef_app
// (src/factor_indef.hxx:1484-1646) with the kernels of src/kernels/ldlt_app.hxx and
// spral/src/ssids/cpu/kernels/{block_ldlt.hxx,ldlt_tpp.cxx}.  Same pivot tests, same
// D^-1 storage, same a-posteriori acceptance rule (|l_ij| <= 1/u); the *organisation*
// is GPU-native:
//   * the block column is IB = 32 wide: one warp factorizes the 32x32 diagonal block with
//     complete pivoting (k_ldlt_diag32  ~ block_ldlt<T,32>),
//   * every row below is one thread's register row: solve with L_bb^T, scale by D^-1 and
//     test the threshold in one pass (k_apply32 ~ apply_pivot<OP_N> + check_threshold),
//   * failed columns are not left in place: after each block column they are swapped
//     (symmetrically) behind the still-active candidates (k_swap_failed), so eliminated
//     columns are always the contiguous prefix [0, p0) and the right-looking update is one
//     dense DMMA rank-k update of everything behind p0 (no ApplyT/UpdateT tasks, no
//     permute_failed at the end, no full-panel backup: the L panel keeps the original
//     entries until a column has passed; W = L*D lives in a scratch panel),
//   * columns that never pass get a second chance with threshold partial pivoting
//     (k_tpp ~ ldlt_tpp_factor) and what is left is delayed to the parent.
#pragma once
#include <cfloat>
#include <cmath>

#include "kernels.cuh"

namespace sylver_b200 {

constexpr int IB = 32;            // inner block / panel width of the APTP path
constexpr int SLD = 33;           // shared-memory stride of the 32x32 block (conflict free)

// ---------------------------------------------------------------------------
// k_ldlt_diag32: one warp per front factorizes the diagonal block
// A[p0:p0+wb, p0:p0+wb] with complete pivoting (block_ldlt.hxx:289-414:
// find_maxloc:77, test_2x2:214, update_1x1:233, update_2x2:222).  Results go to the
// per-front scratch (the L panel is left untouched until columns have passed):
//   S (lower)  = L_bb, unit diagonal implied;   S (upper)[k][r] = (L D)[r][k]
//   dinv[2*j], dinv[2*j+1]  D^-1 in the reference's convention, lperm[j] local pivot order
// ---------------------------------------------------------------------------
struct DiagScratch {
   double S[IB * SLD];
   double dinv[2 * IB + 2];
   int lperm[IB];
   int pad[2];
};

// begin_ob > 0: this is the first block column of an outer panel of up to begin_ob candidates
// (what k_outer_begin did in a launch of its own).
__device__ void ldlt_diag32_front(const DevTree& T, int f, DiagScratch* out, double u, double small, int begin_ob,
                                  double* S, double* dinv, int* lperm) {
   FrontState& st = T.state[f];
   const int lane = threadIdx.x & 31;
   if (begin_ob > 0) {
      if (lane == 0) {
         st.obeg = st.p0;
         st.oend = min(st.p0 + begin_ob, st.na);
         st.pa = st.oend;
      }
      __syncwarp();
   }
   const int p0 = st.p0;
   const int wb = min(IB, st.pa - p0);
   if (lane == 0) st.wb = wb;
   if (wb <= 0) return;
   const int ldl = T.ldl[f];
   const double* A = T.L + T.loff[f] + (size_t)p0 * ldl + p0;
   // load (full symmetric storage)
   // coalesced reads of the lower triangle, mirrored through shared memory
#pragma unroll
   for (int c = 0; c < IB; ++c) {
      double v = 0.0;
      if (lane < wb && c < wb && lane >= c) v = A[(size_t)c * ldl + lane];
      if (lane >= c) {
         S[lane * SLD + c] = v;
         S[c * SLD + lane] = v;
      }
   }
   lperm[lane] = lane;
   dinv[2 * lane] = 0.0;
   dinv[2 * lane + 1] = 0.0;
   if (lane < 2) dinv[2 * IB + lane] = 0.0;
   __syncwarp();
   int p = 0;
   while (p < wb) {
      // ---- largest entry of the remaining lower triangle, first in column-major order ----
      double best = -1.0;
      int bc = IB;
      if (lane >= p && lane < wb) {
         for (int c = p; c <= lane; ++c) {
            const double v = fabs(S[lane * SLD + c]);
            if (v > best) { best = v; bc = c; }
         }
      }
      int br = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
         const double ov = __shfl_xor_sync(0xffffffffu, best, o);
         const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
         const int orr = __shfl_xor_sync(0xffffffffu, br, o);
         const bool take = (ov > best) || (ov == best && (oc < bc || (oc == bc && orr < br)));
         if (take) { best = ov; bc = oc; br = orr; }
      }
      int t = br, m = bc;            // t >= m
      if (best < small) {
         // everything left is negligible: zero pivots (block_ldlt.hxx:303-317)
         for (int q = p; q < wb; ++q) {
            if (lane >= q && lane < IB) { S[lane * SLD + q] = 0.0; S[q * SLD + lane] = 0.0; }
            if (lane == 0) { dinv[2 * q] = 0.0; dinv[2 * q + 1] = 0.0; }
         }
         __syncwarp();
         p = wb;
         break;
      }
      int pivsiz;
      double a11, a21 = 0.0, a22 = 0.0, detscale = 0.0, detpiv = 0.0;
      if (t == m) {
         a11 = S[t * SLD + t];
         pivsiz = 1;
      } else {
         a11 = S[m * SLD + m];
         a22 = S[t * SLD + t];
         a21 = S[t * SLD + m];
         detscale = 1.0 / fabs(a21);
         detpiv = (a11 * detscale) * a22 - fabs(a21);
         if (fabs(detpiv) >= fabs(a21) / 2) {
            pivsiz = 2;
         } else if (fabs(a11) > fabs(a22)) {
            pivsiz = 1; t = m;                 // a11 as 1x1 (complete pivoting guarantees |a11/a21| >= u here)
         } else {
            pivsiz = 1; a11 = a22; m = t;
         }
      }
      __syncwarp();
      // ---- symmetric swaps (rows by lanes-as-columns, columns by lanes-as-rows) ----
      auto sym_swap = [&](int i, int j) {
         if (i == j) return;
         { const double x = S[i * SLD + lane]; S[i * SLD + lane] = S[j * SLD + lane]; S[j * SLD + lane] = x; }
         __syncwarp();
         { const double x = S[lane * SLD + i]; S[lane * SLD + i] = S[lane * SLD + j]; S[lane * SLD + j] = x; }
         if (lane == 0) { const int x = lperm[i]; lperm[i] = lperm[j]; lperm[j] = x; }
         __syncwarp();
      };
      if (pivsiz == 1) {
         sym_swap(p, t);
         const double d11 = 1.0 / a11;
         double l = 0.0;
         if (lane > p && lane < wb) {
            l = S[lane * SLD + p] * d11;      // row p (upper) keeps the unscaled copy = (L D)[.,p]
            S[lane * SLD + p] = l;
         }
         __syncwarp();
         if (lane > p && lane < wb) {
            for (int c = p + 1; c <= lane; ++c) {
               const double v = S[lane * SLD + c] - l * S[p * SLD + c];
               S[lane * SLD + c] = v;
               S[c * SLD + lane] = v;
            }
         }
         if (lane == 0) { dinv[2 * p] = d11; dinv[2 * p + 1] = 0.0; }
         __syncwarp();
      } else {
         // t > m: m >= p, t >= p+1
         sym_swap(p, m);
         if (t == p) t = m;                    // the entry that sat at p moved to m
         sym_swap(p + 1, t);
         const double d11 = (a22 * detscale) / detpiv;
         const double d22 = (a11 * detscale) / detpiv;
         const double d21 = (-a21 * detscale) / detpiv;
         double l1 = 0.0, l2 = 0.0;
         if (lane > p + 1 && lane < wb) {
            const double w1 = S[lane * SLD + p], w2 = S[lane * SLD + p + 1];
            l1 = d11 * w1 + d21 * w2;
            l2 = d21 * w1 + d22 * w2;
            S[lane * SLD + p] = l1;
            S[lane * SLD + p + 1] = l2;
         }
         if (lane == p + 1) S[lane * SLD + p] = 0.0;     // L(p+1,p) = 0; S[p][p+1] keeps a21 = (L D)(p+1,p)
         __syncwarp();
         if (lane > p + 1 && lane < wb) {
            for (int c = p + 2; c <= lane; ++c) {
               const double v = S[lane * SLD + c] - (l1 * S[p * SLD + c] + l2 * S[(p + 1) * SLD + c]);
               S[lane * SLD + c] = v;
               S[c * SLD + lane] = v;
            }
         }
         if (lane == 0) {
            dinv[2 * p] = d11; dinv[2 * p + 1] = d21;
            dinv[2 * p + 2] = INFINITY; dinv[2 * p + 3] = d22;
         }
         __syncwarp();
      }
      p += pivsiz;
   }
   __syncwarp();
   DiagScratch& o = *out;
#pragma unroll
   for (int r = 0; r < IB; ++r) o.S[r * SLD + lane] = S[r * SLD + lane];      // coalesced
   o.dinv[2 * lane] = dinv[2 * lane];
   o.dinv[2 * lane + 1] = dinv[2 * lane + 1];
   if (lane < 2) o.dinv[2 * IB + lane] = 0.0;
   o.lperm[lane] = lperm[lane];
   if (lane == 0) st.npass = wb;    // the apply kernel lowers this with atomicMin
}

// begin_ob > 0: this is the first block column of an outer panel of up to begin_ob candidates
// (what k_outer_begin did in a launch of its own).
static __global__ void __launch_bounds__(32) k_ldlt_diag32(DevTree T, const int* __restrict__ fronts,
                                                           DiagScratch* __restrict__ scratch, double u, double small,
                                                           int begin_ob) {
   __shared__ double S[IB * SLD];
   __shared__ double dinv[2 * IB + 2];
   __shared__ int lperm[IB];
   ldlt_diag32_front(T, fronts[blockIdx.x], scratch + blockIdx.x, u, small, begin_ob, S, dinv, lperm);
}

// K = 32 panel update (k_gemm_batched mode 3) whose first CTA per front also factorizes the NEXT
// diagonal block as soon as it has written it: one launch fewer per block column, and the 45 us
// of the pivoted 32 x 32 factorization overlap with the rest of the panel update.
static __global__ void __launch_bounds__(GT_THREADS, 2)
k_gemm_diag3(DevTree T, TileBatch batch, DiagScratch* __restrict__ scratch, double u, double small) {
   extern __shared__ __align__(128) double smem[];
   gemm_body<true>(T, batch, 3, 0, IB, nullptr, 0, 0, 1,
                   FusedDiag{reinterpret_cast<char*>(scratch), sizeof(DiagScratch), u, small}, smem);
}

// per-column D^-1 application coefficients: l[j] = cs*w[j] + co*w[partner]
struct PivCoef {
   double cs, co;
   int partner;      // j (1x1), j+1 (first of a 2x2) or j-1 (second)
};
__device__ __forceinline__ void pivot_coefs(const double* dinv, int wb, int j, double& cs, double& co, int& kind) {
   // kind 0: 1x1, 1: first of 2x2, 2: second of 2x2, 3: zero pivot
   const bool second = (j > 0) && isinf(dinv[2 * j]);
   if (second) { kind = 2; cs = dinv[2 * j + 1]; co = dinv[2 * j - 1]; return; }
   const bool first = (j + 1 < wb) && isinf(dinv[2 * j + 2]);
   if (first) { kind = 1; cs = dinv[2 * j]; co = dinv[2 * j + 1]; return; }
   cs = dinv[2 * j]; co = 0.0;
   kind = (cs == 0.0) ? 3 : 0;
}

// Work descriptor of the row-parallel kernels: item -> (front, 256-row chunk)
constexpr int AP_THREADS = 256;

// ---------------------------------------------------------------------------
// k_apply32: rows below the diagonal block.  One thread = one row:
//   a' = A[r, p0 + lperm[.]]      (column permutation of the block)
//   w  = a' L_bb^-T               (unit lower solve; w = (L D)[r,.])
//   l  = w D^-1                   (apply_pivot<OP_N>, ldlt_app.hxx:172-214, zero pivots too)
//   first column with |l| > 1/u   (check_threshold, ldlt_app.hxx:144-163) -> atomicMin(npass)
// w goes to the W panel; L is only written once the pass count is final (k_finish32).
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(AP_THREADS) k_apply32(DevTree T, TileBatch batch,
                                                               const DiagScratch* __restrict__ scratch, double u,
                                                               double small) {
   __shared__ double Lb[IB * SLD];
   __shared__ double dinv[2 * IB + 2];
   __shared__ int lperm[IB];
   __shared__ int s_fail;
   const int item = blockIdx.x;
   const int fi = find_front(batch, item);
   const int f = batch.fronts[fi];
   const int chunk = item - batch.prefix[fi];
   const FrontState st = T.state[f];
   const int wb = st.wb;
   if (wb <= 0) return;
   const int p0 = st.p0;
   const int m = T.m[f], ldl = T.ldl[f];
   const int r0 = p0 + wb + chunk * AP_THREADS;
   if (r0 >= m) return;
   const DiagScratch& ds = scratch[fi];
   for (int i = threadIdx.x; i < IB * SLD; i += AP_THREADS) Lb[i] = ds.S[i];
   if (threadIdx.x < 2 * IB + 2) dinv[threadIdx.x] = ds.dinv[threadIdx.x];
   if (threadIdx.x < IB) lperm[threadIdx.x] = ds.lperm[threadIdx.x];
   if (threadIdx.x == 0) s_fail = IB;
   __syncthreads();
   const int r = r0 + threadIdx.x;
   int fail = IB;
   if (r < m) {
      const double* Ar = T.L + T.loff[f] + (size_t)p0 * ldl + r;
      double* Wr = T.W + T.woff[f] + (size_t)p0 * ldl + r;
      double w[IB];
#pragma unroll
      for (int j = 0; j < IB; ++j) w[j] = (j < wb) ? Ar[(size_t)lperm[j] * ldl] : 0.0;
#pragma unroll
      for (int j = 1; j < IB; ++j) {
         double s = w[j];
#pragma unroll
         for (int k = 0; k < j; ++k) s -= w[k] * Lb[j * SLD + k];
         w[j] = s;
      }
      const double lim = 1.0 / u;
#pragma unroll
      for (int j = 0; j < IB; ++j) {
         if (j < wb) {
            double cs, co;
            int kind;
            pivot_coefs(dinv, wb, j, cs, co, kind);
            const double wo = (kind == 1) ? w[(j + 1) & (IB - 1)] : w[(j + IB - 1) & (IB - 1)];
            double l;
            if (kind == 3) l = (fabs(w[j]) < small) ? 0.0 : INFINITY * w[j];
            else l = cs * w[j] + ((kind == 0) ? 0.0 : co * wo);
            if (!(fabs(l) <= lim) && fail == IB) fail = j;      // NaN fails too
            Wr[(size_t)j * ldl] = w[j];
         }
      }
   }
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) fail = min(fail, __shfl_xor_sync(0xffffffffu, fail, o));
   if ((threadIdx.x & 31) == 0 && fail < IB) atomicMin(&s_fail, fail);
   __syncthreads();
   if (threadIdx.x == 0 && s_fail < IB) atomicMin(&T.state[f].npass, s_fail);
}

// number of accepted pivots after Column::adjust (include/sylver/kernels/Column.hxx:87-101):
// never split a 2x2
__device__ __forceinline__ int adjusted_pass(int npass, const double* dinv) {
   if (npass > 0) {
      const double d11 = dinv[2 * (npass - 1)], d21 = dinv[2 * (npass - 1) + 1];
      if (isfinite(d11) && d21 != 0.0) --npass;
   }
   return npass;
}

// ---------------------------------------------------------------------------
// k_finish32: with the pass count final, write the block column in place:
//   passed columns j < npass : L[r, p0+j] = (w D^-1)[j]
//   failed columns j >= npass: the ORIGINAL entries, column-permuted by lperm
//     (what restore_if_required achieves from a backup, ldlt_app.hxx:404-431)
// chunk 0 also writes the diagonal block (L_bb / restored entries), D^-1, perm and
// permutes the rows p0..p0+wb of every column left of the block (apply_rperm).
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(AP_THREADS) k_finish32(DevTree T, TileBatch batch,
                                                                const DiagScratch* __restrict__ scratch,
                                                                double small) {
   __shared__ double Lb[IB * SLD];
   __shared__ double Ao[IB * SLD];
   __shared__ double dinv[2 * IB + 2];
   __shared__ int lperm[IB];
   __shared__ int operm[IB];
   const int item = blockIdx.x;
   const int fi = find_front(batch, item);
   const int f = batch.fronts[fi];
   const int chunk = item - batch.prefix[fi];
   const FrontState st = T.state[f];
   const int wb = st.wb;
   if (wb <= 0) return;
   const int p0 = st.p0;
   const int m = T.m[f], ldl = T.ldl[f];
   const int nchunks = batch.prefix[fi + 1] - batch.prefix[fi];
   // chunk c >= 1 handles rows r0..; chunk 0 handles the diagonal block and the left rows
   const DiagScratch& ds = scratch[fi];
   for (int i = threadIdx.x; i < IB * SLD; i += AP_THREADS) Lb[i] = ds.S[i];
   if (threadIdx.x < 2 * IB + 2) dinv[threadIdx.x] = ds.dinv[threadIdx.x];
   if (threadIdx.x < IB) lperm[threadIdx.x] = ds.lperm[threadIdx.x];
   __syncthreads();
   const int npass = adjusted_pass(min(st.npass, wb), dinv);
   double* Lf = T.L + T.loff[f];
   double* Wf = T.W + T.woff[f];
   if (chunk == 0) {
      // ---- diagonal block ----
      const double* A = Lf + (size_t)p0 * ldl + p0;
      for (int i = threadIdx.x; i < IB * IB; i += AP_THREADS) {
         const int rr = i & (IB - 1), cc = i >> 5;
         double v = 0.0;
         if (rr < wb && cc < wb) v = (rr >= cc) ? A[(size_t)cc * ldl + rr] : A[(size_t)rr * ldl + cc];
         Ao[rr * SLD + cc] = v;
      }
      int* perm = T.perm + T.permoff[f] + p0;
      if (threadIdx.x < wb) operm[threadIdx.x] = perm[threadIdx.x];
      __syncthreads();
      for (int i = threadIdx.x; i < IB * IB; i += AP_THREADS) {
         const int rr = i & (IB - 1), cc = i >> 5;
         if (rr < wb && cc < wb && rr >= cc) {
            double v;
            if (cc < npass) v = (rr == cc) ? 1.0 : Lb[rr * SLD + cc];
            else v = Ao[lperm[rr] * SLD + lperm[cc]];
            Lf[(size_t)(p0 + cc) * ldl + p0 + rr] = v;
            if (cc < npass && rr > cc) Wf[(size_t)(p0 + cc) * ldl + p0 + rr] = Lb[cc * SLD + rr];
         }
      }
      if (threadIdx.x < wb) perm[threadIdx.x] = operm[lperm[threadIdx.x]];
      double* D = T.D + T.doff[f] + 2 * (size_t)p0;
      if (threadIdx.x < 2 * npass) D[threadIdx.x] = dinv[threadIdx.x];
      return;
   }
   // ---- rows p0..p0+wb of the L columns left of the block (apply_rperm): one warp per column,
   // one lane per row (a 256 B segment read and written once), shared between the row chunks.
   // Skipped when the block needed no permutation.  The W panel is NOT permuted: its rows in the
   // candidate range are dead once the block column's own trailing update has been applied
   // (later updates read only their own block's W columns, the contribution update rows >= n).
   {
      const int lane = threadIdx.x & 31;
      const int src = (lane < wb) ? lperm[lane] : lane;
      if (__any_sync(0xffffffffu, src != lane)) {
         const int nw = AP_THREADS / 32;
         for (int c = (chunk - 1) * nw + (threadIdx.x >> 5); c < p0; c += (nchunks - 1) * nw) {
            double* col = Lf + (size_t)c * ldl + p0;
            const double v = (lane < wb) ? col[src] : 0.0;
            __syncwarp();
            if (lane < wb) col[lane] = v;
         }
      }
   }
   const int r = p0 + wb + (chunk - 1) * AP_THREADS + threadIdx.x;
   if (r >= m) return;
   double* Ar = Lf + (size_t)p0 * ldl + r;
   const double* Wr = Wf + (size_t)p0 * ldl + r;
   double out[IB];
#pragma unroll
   for (int j = 0; j < IB; ++j) {
      if (j < npass) {
         double cs, co;
         int kind;
         pivot_coefs(dinv, wb, j, cs, co, kind);
         const double wj = Wr[(size_t)j * ldl];
         if (kind == 3) out[j] = (fabs(wj) < small) ? 0.0 : INFINITY * wj;
         else if (kind == 0) out[j] = cs * wj;
         else out[j] = cs * wj + co * Wr[(size_t)(kind == 1 ? j + 1 : j - 1) * ldl];
      } else if (j < wb) {
         out[j] = Ar[(size_t)lperm[j] * ldl];
      }
   }
#pragma unroll
   for (int j = 0; j < IB; ++j)
      if (j < wb) Ar[(size_t)j * ldl] = out[j];
}

// ---------------------------------------------------------------------------
// symmetric swap of candidate variables i < j of a front whose eliminated prefix is
// `nleft` columns wide: lower-triangle panel L (all columns < n) and the rows of W in
// the eliminated columns (swap_cols, ldlt_tpp.cxx:44-72).  Called by a whole CTA.
// ---------------------------------------------------------------------------
// W rows are only swapped in the columns [wfirst, nleft): the pivots whose trailing update is
// still to come (older W columns are dead in the candidate row range).
__device__ __forceinline__ void cta_sym_swap(double* Lf, double* Wf, int* perm, int ldl, int m, int n, int nleft,
                                             int i, int j, int wfirst) {
   if (i == j) return;
   if (i > j) { const int x = i; i = j; j = x; }
   const int tid = threadIdx.x, nt = blockDim.x;
   // rows i and j in columns c < i (L) and c < nleft (W)
   for (int c = tid; c < i; c += nt) {
      double* col = Lf + (size_t)c * ldl;
      const double x = col[i]; col[i] = col[j]; col[j] = x;
      if (c < nleft && c >= wfirst) {
         double* wc = Wf + (size_t)c * ldl;
         const double y = wc[i]; wc[i] = wc[j]; wc[j] = y;
      }
   }
   // a(i+1:j-1, i)  <->  a(j, i+1:j-1)
   for (int k = i + 1 + tid; k < j; k += nt) {
      double* p1 = Lf + (size_t)i * ldl + k;
      double* p2 = Lf + (size_t)k * ldl + j;
      const double x = *p1; *p1 = *p2; *p2 = x;
   }
   // a(j+1:m, i) <-> a(j+1:m, j)   (column j only exists if j < n)
   if (j < n) {
      for (int r = j + 1 + tid; r < m; r += nt) {
         double* p1 = Lf + (size_t)i * ldl + r;
         double* p2 = Lf + (size_t)j * ldl + r;
         const double x = *p1; *p1 = *p2; *p2 = x;
      }
   }
   if (tid == 0) {
      double* p1 = Lf + (size_t)i * ldl + i;
      double* p2 = Lf + (size_t)j * ldl + j;
      const double x = *p1; *p1 = *p2; *p2 = x;
      const int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
   }
   __syncthreads();
}

// ---------------------------------------------------------------------------
// k_swap_failed: one CTA per front.  Applies Column::adjust, moves the failed columns of
// the block behind the active candidates and advances the front state:
//   p0 += npass ; na -= nfail ; (kbeg, klen) = the pivots the trailing update must apply.
// ---------------------------------------------------------------------------
constexpr int SW_THREADS = 512;
static __global__ void __launch_bounds__(SW_THREADS) k_swap_failed(DevTree T, const int* __restrict__ fronts,
                                                                   const DiagScratch* __restrict__ scratch) {
   const int f = fronts[blockIdx.x];
   FrontState& st = T.state[f];
   const int wb = st.wb;
   if (wb <= 0) {
      if (threadIdx.x == 0) st.klen = 0;
      return;
   }
   const int p0 = st.p0, pa = st.pa;
   const int npass = adjusted_pass(min(st.npass, wb), scratch[blockIdx.x].dinv);
   const int nf = wb - npass;
   const int q = pa - p0 - wb;
   const int ns = min(nf, q);
   __syncthreads();
   if (ns > 0) {
      const int m = T.m[f], n = T.n[f], ldl = T.ldl[f];
      double* Lf = T.L + T.loff[f];
      double* Wf = T.W + T.woff[f];
      int* perm = T.perm + T.permoff[f];
      for (int i = 0; i < ns; ++i) cta_sym_swap(Lf, Wf, perm, ldl, m, n, p0 + npass, p0 + npass + i, pa - ns + i, p0);
   }
   if (threadIdx.x == 0) {
      st.kbeg = p0;
      st.klen = npass;
      st.p0 = p0 + npass;
      st.pa = pa - nf;       // failed columns stay inside the outer panel (they keep receiving its updates)
      st.npass = IB;
      st.wb = 0;
   }
}

// ---------------------------------------------------------------------------
// k_block_column32: k_apply32 + k_finish32 + k_swap_failed in ONE launch.
//   phase A  every CTA of the front applies the pivots to its rows (w kept in registers) and
//            lowers npass; the row permutation of the columns left of the block is done here too
//   phase B  front-wide rendezvous: the CTAs of a front have consecutive block indices and are
//            dispatched in order, so the ones that arrived can wait for the rest (a front has at
//            most 1 + m/256 of them, far fewer than fit on the GPU at once)
//   phase C  with the pass count final, every CTA writes its rows in place (k_finish32)
//   phase D  the LAST CTA to finish moves the failed columns behind the active candidates and
//            advances the front state (k_swap_failed)
// One thread = one row; chunk 0 = the diagonal block.
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(AP_THREADS) k_block_column32(DevTree T, TileBatch batch,
                                                                      const DiagScratch* __restrict__ scratch, double u,
                                                                      double small) {
   __shared__ double Lb[IB * SLD];
   __shared__ double Ao[IB * SLD];
   __shared__ double dinv[2 * IB + 2];
   __shared__ int lperm[IB];
   __shared__ int operm[IB];
   __shared__ int s_fail, s_npass, s_last;
   const int item = blockIdx.x;
   const int fi = find_front(batch, item);
   const int f = batch.fronts[fi];
   const int chunk = item - batch.prefix[fi];
   FrontState& gst = T.state[f];
   const FrontState st = gst;
   const int wb = st.wb;
   if (wb <= 0) return;
   const int p0 = st.p0;
   const int m = T.m[f], n = T.n[f], ldl = T.ldl[f];
   // at least one row chunk even when no row lies below the block (root fronts): it still has
   // to permute the rows of the columns left of the block
   const int nrowchunks = max((m - (p0 + wb) + AP_THREADS - 1) / AP_THREADS, 1);
   const int nexp = 1 + nrowchunks;                    // CTAs of this front that take part
   if (chunk >= nexp) return;
   const int nchunks = batch.prefix[fi + 1] - batch.prefix[fi];
   const DiagScratch& ds = scratch[fi];
   for (int i = threadIdx.x; i < IB * SLD; i += AP_THREADS) Lb[i] = ds.S[i];
   if (threadIdx.x < 2 * IB + 2) dinv[threadIdx.x] = ds.dinv[threadIdx.x];
   if (threadIdx.x < IB) lperm[threadIdx.x] = ds.lperm[threadIdx.x];
   if (threadIdx.x == 0) s_fail = IB;
   __syncthreads();
   double* Lf = T.L + T.loff[f];
   double* Wf = T.W + T.woff[f];
   const int r = p0 + wb + (chunk - 1) * AP_THREADS + threadIdx.x;
   const bool rowthread = chunk >= 1 && r < m;
   double w[IB];
   // ---------------- phase A ----------------
   if (chunk == 0) {
      const double* A = Lf + (size_t)p0 * ldl + p0;
      for (int i = threadIdx.x; i < IB * IB; i += AP_THREADS) {
         const int rr = i & (IB - 1), cc = i >> 5;
         double v = 0.0;
         if (rr < wb && cc < wb) v = (rr >= cc) ? A[(size_t)cc * ldl + rr] : A[(size_t)rr * ldl + cc];
         Ao[rr * SLD + cc] = v;
      }
      if (threadIdx.x < wb) operm[threadIdx.x] = T.perm[T.permoff[f] + p0 + threadIdx.x];
   } else {
      // rows p0..p0+wb of the L columns left of the block (apply_rperm): one warp per column, one
      // lane per row, shared between the row chunks; skipped when the block did not permute.  The W
      // panel is NOT permuted (its candidate rows are dead after the block column's own update).
      const int lane = threadIdx.x & 31;
      const int src = (lane < wb) ? lperm[lane] : lane;
      if (__any_sync(0xffffffffu, src != lane)) {
         const int nw = AP_THREADS / 32;
         for (int c = (chunk - 1) * nw + (threadIdx.x >> 5); c < p0; c += nrowchunks * nw) {
            double* col = Lf + (size_t)c * ldl + p0;
            const double v = (lane < wb) ? col[src] : 0.0;
            __syncwarp();
            if (lane < wb) col[lane] = v;
         }
      }
      int fail = IB;
      if (rowthread) {
         const double* Ar = Lf + (size_t)p0 * ldl + r;
         double* Wr = Wf + (size_t)p0 * ldl + r;
#pragma unroll
         for (int j = 0; j < IB; ++j) w[j] = (j < wb) ? Ar[(size_t)lperm[j] * ldl] : 0.0;
#pragma unroll
         for (int j = 1; j < IB; ++j) {
            double s = w[j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= w[k] * Lb[j * SLD + k];
            w[j] = s;
         }
         const double lim = 1.0 / u;
#pragma unroll
         for (int j = 0; j < IB; ++j) {
            if (j < wb) {
               double cs, co;
               int kind;
               pivot_coefs(dinv, wb, j, cs, co, kind);
               const double wo = (kind == 1) ? w[(j + 1) & (IB - 1)] : w[(j + IB - 1) & (IB - 1)];
               double l;
               if (kind == 3) l = (fabs(w[j]) < small) ? 0.0 : INFINITY * w[j];
               else l = cs * w[j] + ((kind == 0) ? 0.0 : co * wo);
               if (!(fabs(l) <= lim) && fail == IB) fail = j;      // NaN fails too
               Wr[(size_t)j * ldl] = w[j];
            }
         }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) fail = min(fail, __shfl_xor_sync(0xffffffffu, fail, o));
      if ((threadIdx.x & 31) == 0 && fail < IB) atomicMin(&s_fail, fail);
   }
   __syncthreads();
   // ---------------- phase B: rendezvous ----------------
   if (threadIdx.x == 0) {
      if (s_fail < IB) atomicMin(&gst.npass, s_fail);
      __threadfence();
      atomicAdd(&gst.arrive, 1);
      while (atomicAdd(&gst.arrive, 0) < nexp) __nanosleep(64);
      __threadfence();
      s_npass = atomicAdd(&gst.npass, 0);
   }
   __syncthreads();
   // ---------------- phase C: finish ----------------
   const int npass = adjusted_pass(min(s_npass, wb), dinv);
   if (chunk == 0) {
      for (int i = threadIdx.x; i < IB * IB; i += AP_THREADS) {
         const int rr = i & (IB - 1), cc = i >> 5;
         if (rr < wb && cc < wb && rr >= cc) {
            double v;
            if (cc < npass) v = (rr == cc) ? 1.0 : Lb[rr * SLD + cc];
            else v = Ao[lperm[rr] * SLD + lperm[cc]];
            Lf[(size_t)(p0 + cc) * ldl + p0 + rr] = v;
            if (cc < npass && rr > cc) Wf[(size_t)(p0 + cc) * ldl + p0 + rr] = Lb[cc * SLD + rr];
         }
      }
      int* perm = T.perm + T.permoff[f] + p0;
      if (threadIdx.x < wb) perm[threadIdx.x] = operm[lperm[threadIdx.x]];
      double* D = T.D + T.doff[f] + 2 * (size_t)p0;
      if (threadIdx.x < 2 * npass) D[threadIdx.x] = dinv[threadIdx.x];
   } else if (rowthread) {
      double* Ar = Lf + (size_t)p0 * ldl + r;
      double out[IB];
#pragma unroll
      for (int j = 0; j < IB; ++j) {
         if (j < npass) {
            double cs, co;
            int kind;
            pivot_coefs(dinv, wb, j, cs, co, kind);
            const double wo = (kind == 1) ? w[(j + 1) & (IB - 1)] : w[(j + IB - 1) & (IB - 1)];
            if (kind == 3) out[j] = (fabs(w[j]) < small) ? 0.0 : INFINITY * w[j];
            else if (kind == 0) out[j] = cs * w[j];
            else out[j] = cs * w[j] + co * wo;
         } else if (j < wb) {
            out[j] = Ar[(size_t)lperm[j] * ldl];      // failed: the original entry, column-permuted
         }
      }
#pragma unroll
      for (int j = 0; j < IB; ++j)
         if (j < wb) Ar[(size_t)j * ldl] = out[j];
   }
   // ---------------- phase D: the last CTA advances the front ----------------
   __syncthreads();
   if (threadIdx.x == 0) {
      __threadfence();
      s_last = (atomicAdd(&gst.done, 1) == nexp - 1) ? 1 : 0;
   }
   __syncthreads();
   if (!s_last) return;
   __threadfence();
   {
      const int pa = st.pa;
      const int nf = wb - npass;
      const int q = pa - p0 - wb;
      const int ns = min(nf, q);
      if (ns > 0) {
         int* perm = T.perm + T.permoff[f];
         for (int i = 0; i < ns; ++i) cta_sym_swap(Lf, Wf, perm, ldl, m, n, p0 + npass, p0 + npass + i, pa - ns + i, p0);
      }
      if (threadIdx.x == 0) {
         gst.kbeg = p0;
         gst.klen = npass;
         gst.p0 = p0 + npass;
         gst.pa = pa - nf;       // failed columns stay inside the outer panel (they keep receiving its updates)
         gst.npass = IB;
         gst.wb = 0;
         gst.arrive = 0;
         gst.done = 0;
      }
   }
   (void)nchunks;
}

// Outer panel of the two-level blocking: up to OB candidates.
constexpr int OB_DEFAULT = 256;      // SYLVER_B200_OB overrides (multiple of 32)
static __global__ void k_outer_begin(DevTree T, const int* __restrict__ fronts, int cnt, int OB) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= cnt) return;
   FrontState& st = T.state[fronts[i]];
   st.obeg = st.p0;
   st.oend = min(st.p0 + OB, st.na);
   st.pa = st.oend;
}
// After the panel's rank-(p0 - obeg) update of the columns behind it, every column >= p0 has
// seen the same pivots: the panel's failed columns [p0, oend) move behind the still-active
// candidates (one CTA per front).
// `slot` (panel parity): the panel's pivots [obeg, p0) are published in cb/ce[slot] for its
// contribution-block pass (k_gemm_batched mode 6).
static __global__ void __launch_bounds__(SW_THREADS) k_outer_end(DevTree T, const int* __restrict__ fronts, int slot) {
   const int f = fronts[blockIdx.x];
   FrontState& st = T.state[f];
   const int p0 = st.p0, na = st.na, oend = st.oend;
   const int nf = oend - p0;
   if (threadIdx.x == 0) {
      st.cb[slot] = st.obeg;
      st.ce[slot] = p0;
      st.co[slot] = oend;
      st.cf[slot] = min(nf, na - oend) > 0 ? 1 : 0;
   }
   if (nf <= 0) return;
   const int ns = min(nf, na - oend);
   __syncthreads();
   if (ns > 0) {
      const int m = T.m[f], n = T.n[f], ldl = T.ldl[f];
      double* Lf = T.L + T.loff[f];
      double* Wf = T.W + T.woff[f];
      int* perm = T.perm + T.permoff[f];
      for (int i = 0; i < ns; ++i) cta_sym_swap(Lf, Wf, perm, ldl, m, n, p0, p0 + i, na - ns + i, p0);
   }
   if (threadIdx.x == 0) {
      st.na = na - nf;
      st.oend = p0;
      st.pa = p0;
   }
}

// ---------------------------------------------------------------------------
// k_tpp: second pass on the columns that failed APTP -- threshold partial pivoting,
// one CTA per front, statement by statement after ldlt_tpp_factor
// (spral/src/ssids/cpu/kernels/ldlt_tpp.cxx:166-270) on the trailing
// (m-p0) x (n-p0) panel; "aleft" are the rows of the eliminated L and W columns.
// Writes L, W = L*D (the reference's ld workspace) and D^-1.
// ---------------------------------------------------------------------------
constexpr int TPP_THREADS = 512;

__device__ __forceinline__ double cta_max(double v, double* red) {
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
   __syncthreads();
   if (lane == 0) red[warp] = v;
   __syncthreads();
   double r = red[0];
   for (int i = 1; i < TPP_THREADS / 32; ++i) r = fmax(r, red[i]);
   return r;
}
// largest |a(col, c)|, nelim <= c < col, and |a(r, col)|, col < r < m, excluding index `ex`
__device__ __forceinline__ double tpp_rc_max(const double* Lf, int ldl, int col, int nelim, int m, int ex, double* red) {
   double best = 0.0;
   for (int c = nelim + threadIdx.x; c < col; c += TPP_THREADS)
      if (c != ex) best = fmax(best, fabs(Lf[(size_t)c * ldl + col]));
   for (int r = col + 1 + threadIdx.x; r < m; r += TPP_THREADS)
      if (r != ex) best = fmax(best, fabs(Lf[(size_t)col * ldl + r]));
   return cta_max(best, red);
}

static __global__ void __launch_bounds__(TPP_THREADS) k_tpp(DevTree T, const int* __restrict__ fronts, double u,
                                                            double small, int force_root_only) {
   __shared__ double red[TPP_THREADS / 32];
   __shared__ int s_idx;
   const int f = fronts[blockIdx.x];
   FrontState& st = T.state[f];
   const int m = T.m[f], n = T.n[f], ldl = T.ldl[f];
   int nelim = st.p0;
   if (threadIdx.x == 0) st.nelim1 = nelim;
   if (nelim >= n || (force_root_only && m != n)) {
      if (threadIdx.x == 0) st.nelim = nelim;
      return;
   }
   double* Lf = T.L + T.loff[f];
   double* Wf = T.W + T.woff[f];
   double* D = T.D + T.doff[f];
   int* perm = T.perm + T.permoff[f];
   const int tid = threadIdx.x;
   auto col_small = [&](int idx) -> bool {      // check_col_small(idx, nelim, m)
      double best = 0.0;
      for (int c = nelim + tid; c < idx; c += TPP_THREADS) best = fmax(best, fabs(Lf[(size_t)c * ldl + idx]));
      for (int r = idx + tid; r < m; r += TPP_THREADS) best = fmax(best, fabs(Lf[(size_t)idx * ldl + r]));
      return cta_max(best, red) < small;
   };
   auto zero_pivot = [&]() {
      for (int r = nelim + tid; r < m; r += TPP_THREADS) {
         Lf[(size_t)nelim * ldl + r] = 0.0;
         Wf[(size_t)nelim * ldl + r] = 0.0;
      }
      if (tid == 0) { D[2 * nelim] = 0.0; D[2 * nelim + 1] = 0.0; }
      __syncthreads();
      ++nelim;
   };
   auto elim_1x1 = [&]() {      // pivot sits at (nelim, nelim)
      const double d11 = 1.0 / Lf[(size_t)nelim * ldl + nelim];
      __syncthreads();
      double* a1 = Lf + (size_t)nelim * ldl;
      double* w1 = Wf + (size_t)nelim * ldl;
      for (int r = nelim + 1 + tid; r < m; r += TPP_THREADS) {
         const double v = a1[r];
         w1[r] = v;
         a1[r] = v * d11;
      }
      if (tid == 0) { a1[nelim] = 1.0; D[2 * nelim] = d11; D[2 * nelim + 1] = 0.0; }
      __syncthreads();
      // trailing update of the remaining candidate columns
      const int nc = n - nelim - 1;
      for (int c = 0; c < nc; ++c) {
         const int cc = nelim + 1 + c;
         const double wc = w1[cc];
         double* ac = Lf + (size_t)cc * ldl;
         for (int r = cc + tid; r < m; r += TPP_THREADS) ac[r] -= a1[r] * wc;
      }
      __syncthreads();
      nelim += 1;
   };
   auto elim_2x2 = [&](double d11, double d21, double d22) {
      double* a1 = Lf + (size_t)nelim * ldl;
      double* a2 = Lf + (size_t)(nelim + 1) * ldl;
      double* w1 = Wf + (size_t)nelim * ldl;
      double* w2 = Wf + (size_t)(nelim + 1) * ldl;
      for (int r = nelim + 2 + tid; r < m; r += TPP_THREADS) {
         const double v1 = a1[r], v2 = a2[r];
         w1[r] = v1; w2[r] = v2;
         a1[r] = d11 * v1 + d21 * v2;
         a2[r] = d21 * v1 + d22 * v2;
      }
      if (tid == 0) {
         // (L D) rows of the pivot block itself: D = inverse of the stored 2x2
         w1[nelim + 1] = a1[nelim + 1];
         a1[nelim] = 1.0; a1[nelim + 1] = 0.0; a2[nelim + 1] = 1.0;
         D[2 * nelim] = d11; D[2 * nelim + 1] = d21; D[2 * nelim + 2] = INFINITY; D[2 * nelim + 3] = d22;
      }
      __syncthreads();
      const int nc = n - nelim - 2;
      for (int c = 0; c < nc; ++c) {
         const int cc = nelim + 2 + c;
         const double wc1 = w1[cc], wc2 = w2[cc];
         double* ac = Lf + (size_t)cc * ldl;
         for (int r = cc + tid; r < m; r += TPP_THREADS) ac[r] -= a1[r] * wc1 + a2[r] * wc2;
      }
      __syncthreads();
      nelim += 2;
   };
   while (nelim < n) {
      if (col_small(nelim)) { zero_pivot(); continue; }
      int p;
      bool done = false;
      for (p = nelim + 1; p < n; ++p) {
         if (col_small(p)) {
            cta_sym_swap(Lf, Wf, perm, ldl, m, n, nelim, nelim, p, nelim);
            zero_pivot();
            done = true;
            break;
         }
         // t = argmax |a(p, nelim:p-1)| (first maximum)
         {
            double bv = -1.0; int bi = p;
            for (int c = nelim + tid; c < p; c += TPP_THREADS) {
               const double v = fabs(Lf[(size_t)c * ldl + p]);
               if (v > bv) { bv = v; bi = c; }
            }
            const double mx = cta_max(bv, red);
            if (tid == 0) s_idx = n;
            __syncthreads();
            if (bv == mx) atomicMin(&s_idx, bi);
            __syncthreads();
         }
         const int t = s_idx;
         const double maxt = tpp_rc_max(Lf, ldl, t, nelim, m, p, red);
         double maxp = tpp_rc_max(Lf, ldl, p, nelim, m, t, red);
         // test_2x2 (ldlt_tpp.cxx:89-119)
         const double a11 = Lf[(size_t)t * ldl + t], a21 = Lf[(size_t)t * ldl + p], a22 = Lf[(size_t)p * ldl + p];
         bool ok2 = false;
         double d11 = 0, d21 = 0, d22 = 0;
         const double maxpiv = fmax(fabs(a11), fmax(fabs(a21), fabs(a22)));
         if (maxpiv >= small) {
            const double detscale = 1 / maxpiv;
            const double detpiv0 = (a11 * detscale) * a22;
            const double detpiv1 = (a21 * detscale) * a21;
            const double detpiv = detpiv0 - detpiv1;
            if (!(fabs(detpiv) < fmax(small, fmax(fabs(detpiv0 / 2), fabs(detpiv1 / 2))))) {
               d11 = (a22 * detscale) / detpiv;
               d21 = (-a21 * detscale) / detpiv;
               d22 = (a11 * detscale) / detpiv;
               if (fmax(maxp, maxt) < small) ok2 = true;
               else {
                  const double x1 = fabs(d11) * maxt + fabs(d21) * maxp;
                  const double x2 = fabs(d21) * maxt + fabs(d22) * maxp;
                  ok2 = (u * fmax(x1, x2) < 1.0);
               }
            }
         }
         __syncthreads();
         if (ok2) {
            cta_sym_swap(Lf, Wf, perm, ldl, m, n, nelim, nelim, t, nelim);
            cta_sym_swap(Lf, Wf, perm, ldl, m, n, nelim, nelim + 1, p, nelim);
            elim_2x2(d11, d21, d22);
            done = true;
            break;
         }
         maxp = fmax(maxp, fabs(a21));
         if (fabs(a22) >= u * maxp) {
            cta_sym_swap(Lf, Wf, perm, ldl, m, n, nelim, nelim, p, nelim);
            elim_1x1();
            done = true;
            break;
         }
      }
      if (!done) {
         // last resort: 1x1 on column nelim
         const double maxp = tpp_rc_max(Lf, ldl, nelim, nelim, m, -1, red);
         if (fabs(Lf[(size_t)nelim * ldl + nelim]) >= u * maxp) {
            __syncthreads();
            elim_1x1();
         } else {
            break;       // out of pivots: the rest is delayed
         }
      }
   }
   if (tid == 0) { st.nelim = nelim; st.p0 = nelim; }
}

// ---------------------------------------------------------------------------
// k_tpp_multi: the same second pass with G CTAs per front.  The pivot search is sequential, but
// every step of it is a reduction over, or an update of, all rows of the front: a front with a
// few thousand failed columns spends O(m F^2) flops here, far too much for one SM.  The G CTAs
// of a front take rows (and, in the row-wise scans, columns) g, g + G, ... in chunks of the CTA
// width, exchange partial maxima through a small per-front scratch and meet at a front-wide
// barrier (same dispatch-order argument as k_block_column32: G <= 16 consecutive CTAs).  All CTAs
// evaluate the decisions redundantly on the same values, so they take the same branches; the
// tests and the arithmetic are those of k_tpp (ldlt_tpp_factor).
// The search itself is batched: the reference examines the candidates p = nelim+1, nelim+2, ...
// one after the other until one can be pivoted, and a rejected candidate changes nothing -- so
// the G CTAs examine G consecutive candidates at once, one each, and the smallest p that can be
// pivoted wins: the same pivot sequence, found in 1/G of the rounds.
// ---------------------------------------------------------------------------
constexpr int TPP_GMAX = 64;    // CTAs per front at most
struct TppScratch {
   double part[2][TPP_GMAX][5];      // [round parity][CTA][value]: partial maxima / a candidate's verdict
};
static_assert(sizeof(TppScratch) <= sizeof(DiagScratch), "TPP scratch lives in the diagonal-block scratch");

struct TppGroup {
   FrontState* st;
   TppScratch* sc;
   int g, G;
   int epoch;                  // barriers passed so far
   int round;                  // reduction rounds so far
};

__device__ __forceinline__ void tpp_barrier(TppGroup& tg) {
   __syncthreads();
   ++tg.epoch;
   if (tg.G > 1 && threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&tg.st->arrive, 1);
      while (atomicAdd(&tg.st->arrive, 0) < tg.G * tg.epoch) __nanosleep(32);
      __threadfence();
   }
   __syncthreads();
}

// CTA-wide max of up to 4 values per thread, then across the front's CTAs; result in v[0..nv)
template <int NV>
__device__ __forceinline__ void tpp_allmax(TppGroup& tg, double (&v)[NV], double* red) {
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
   for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] = fmax(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
   __syncthreads();
   if (lane == 0)
#pragma unroll
      for (int k = 0; k < NV; ++k) red[warp * NV + k] = v[k];
   __syncthreads();
#pragma unroll
   for (int k = 0; k < NV; ++k) {
      double r = red[k];
      for (int i = 1; i < TPP_THREADS / 32; ++i) r = fmax(r, red[i * NV + k]);
      v[k] = r;
   }
   if (tg.G == 1) return;
   const int par = tg.round & 1;
   ++tg.round;
   if (threadIdx.x == 0)
#pragma unroll
      for (int k = 0; k < NV; ++k) tg.sc->part[par][tg.g][k] = v[k];
   tpp_barrier(tg);
#pragma unroll
   for (int k = 0; k < NV; ++k) {
      double r = 0.0;
      for (int i = 0; i < tg.G; ++i) r = fmax(r, __ldcg(&tg.sc->part[par][i][k]));
      v[k] = r;
   }
}

// distributed symmetric swap of candidate variables i < j (cta_sym_swap split over the group)
__device__ __forceinline__ void tpp_sym_swap(TppGroup& tg, double* Lf, double* Wf, int* perm, int ldl, int m, int n,
                                             int nleft, int i, int j, int wfirst) {
   if (i == j) return;
   if (i > j) { const int x = i; i = j; j = x; }
   const int t0 = tg.g * TPP_THREADS + threadIdx.x, nt = tg.G * TPP_THREADS;
   for (int c = t0; c < i; c += nt) {
      double* col = Lf + (size_t)c * ldl;
      const double x = __ldcg(col + i); col[i] = __ldcg(col + j); col[j] = x;
      if (c < nleft && c >= wfirst) {
         double* wc = Wf + (size_t)c * ldl;
         const double y = __ldcg(wc + i); wc[i] = __ldcg(wc + j); wc[j] = y;
      }
   }
   for (int k = i + 1 + t0; k < j; k += nt) {
      double* p1 = Lf + (size_t)i * ldl + k;
      double* p2 = Lf + (size_t)k * ldl + j;
      const double x = __ldcg(p1); *p1 = __ldcg(p2); *p2 = x;
   }
   if (j < n) {
      for (int r = j + 1 + t0; r < m; r += nt) {
         double* p1 = Lf + (size_t)i * ldl + r;
         double* p2 = Lf + (size_t)j * ldl + r;
         const double x = __ldcg(p1); *p1 = __ldcg(p2); *p2 = x;
      }
   }
   if (t0 == 0) {
      double* p1 = Lf + (size_t)i * ldl + i;
      double* p2 = Lf + (size_t)j * ldl + j;
      const double x = __ldcg(p1); *p1 = __ldcg(p2); *p2 = x;
      const int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
   }
   tpp_barrier(tg);
}

static __global__ void __launch_bounds__(TPP_THREADS) k_tpp_multi(DevTree T, const int* __restrict__ fronts,
                                                                  DiagScratch* __restrict__ scratch, double u, double small,
                                                                  int force_root_only, int G) {
   __shared__ double red[(TPP_THREADS / 32) * 4];
   __shared__ int s_idx;
   const int fi = blockIdx.x / G;
   const int f = fronts[fi];
   FrontState& st = T.state[f];
   const int m = T.m[f], n = T.n[f], ldl = T.ldl[f];
   int nelim = st.p0;
   TppGroup tg{&st, reinterpret_cast<TppScratch*>(&scratch[fi]), (int)(blockIdx.x % G), G, 0, 0};
   const bool skip = nelim >= n || (force_root_only && m != n);
   if (skip) {
      // every CTA of the front takes this branch (same state); only one writes
      if (tg.g == 0 && threadIdx.x == 0) { st.nelim1 = nelim; st.nelim = nelim; }
      return;
   }
   double* Lf = T.L + T.loff[f];
   double* Wf = T.W + T.woff[f];
   double* D = T.D + T.doff[f];
   int* perm = T.perm + T.permoff[f];
   const int t0 = tg.g * TPP_THREADS + threadIdx.x, nt = G * TPP_THREADS;
   const int nelim_in = nelim;
   // max |a(idx, c)|, nelim <= c < idx (row part) and |a(r, idx)|, r0 <= r < m (column part), without `ex`
   auto scan = [&](int idx, int r0, int ex) -> double {
      double best = 0.0;
      for (int c = nelim + t0; c < idx; c += nt)
         if (c != ex) best = fmax(best, fabs(__ldcg(Lf + (size_t)c * ldl + idx)));
      for (int r = r0 + t0; r < m; r += nt)
         if (r != ex) best = fmax(best, fabs(__ldcg(Lf + (size_t)idx * ldl + r)));
      return best;
   };
   // the same scan by this CTA alone (its candidate's trial)
   auto scan1 = [&](int idx, int r0, int ex) -> double {
      double best = 0.0;
      for (int c = nelim + threadIdx.x; c < idx; c += TPP_THREADS)
         if (c != ex) best = fmax(best, fabs(__ldcg(Lf + (size_t)c * ldl + idx)));
      for (int r = r0 + threadIdx.x; r < m; r += TPP_THREADS)
         if (r != ex) best = fmax(best, fabs(__ldcg(Lf + (size_t)idx * ldl + r)));
      return best;
   };
   auto cta_max2 = [&](double (&v)[2], double* rd) {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) v[k] = fmax(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
      __syncthreads();
      if (lane == 0) { rd[warp * 2] = v[0]; rd[warp * 2 + 1] = v[1]; }
      __syncthreads();
      double r0 = rd[0], r1 = rd[1];
      for (int i = 1; i < TPP_THREADS / 32; ++i) { r0 = fmax(r0, rd[2 * i]); r1 = fmax(r1, rd[2 * i + 1]); }
      v[0] = r0; v[1] = r1;
   };
   auto zero_pivot = [&]() {
      for (int r = nelim + t0; r < m; r += nt) {
         Lf[(size_t)nelim * ldl + r] = 0.0;
         Wf[(size_t)nelim * ldl + r] = 0.0;
      }
      if (t0 == 0) { D[2 * nelim] = 0.0; D[2 * nelim + 1] = 0.0; }
      tpp_barrier(tg);
      ++nelim;
   };
   auto elim_1x1 = [&]() {      // pivot sits at (nelim, nelim)
      const double d11 = 1.0 / __ldcg(Lf + (size_t)nelim * ldl + nelim);
      tpp_barrier(tg);          // everyone has read the diagonal before it becomes 1
      double* a1 = Lf + (size_t)nelim * ldl;
      double* w1 = Wf + (size_t)nelim * ldl;
      for (int r = nelim + 1 + t0; r < m; r += nt) {
         const double v = __ldcg(a1 + r);
         w1[r] = v;
         a1[r] = v * d11;
      }
      if (t0 == 0) { a1[nelim] = 1.0; D[2 * nelim] = d11; D[2 * nelim + 1] = 0.0; }
      tpp_barrier(tg);
      const int nc = n - nelim - 1;
      for (int c = 0; c < nc; ++c) {
         const int cc = nelim + 1 + c;
         const double wc = __ldcg(w1 + cc);
         double* ac = Lf + (size_t)cc * ldl;
         for (int r = cc + t0; r < m; r += nt) ac[r] = __ldcg(ac + r) - __ldcg(a1 + r) * wc;
      }
      tpp_barrier(tg);
      nelim += 1;
   };
   auto elim_2x2 = [&](double d11, double d21, double d22) {
      double* a1 = Lf + (size_t)nelim * ldl;
      double* a2 = Lf + (size_t)(nelim + 1) * ldl;
      double* w1 = Wf + (size_t)nelim * ldl;
      double* w2 = Wf + (size_t)(nelim + 1) * ldl;
      for (int r = nelim + 2 + t0; r < m; r += nt) {
         const double v1 = __ldcg(a1 + r), v2 = __ldcg(a2 + r);
         w1[r] = v1; w2[r] = v2;
         a1[r] = d11 * v1 + d21 * v2;
         a2[r] = d21 * v1 + d22 * v2;
      }
      if (t0 == 0) {
         w1[nelim + 1] = __ldcg(a1 + nelim + 1);
         a1[nelim] = 1.0; a1[nelim + 1] = 0.0; a2[nelim + 1] = 1.0;
         D[2 * nelim] = d11; D[2 * nelim + 1] = d21; D[2 * nelim + 2] = INFINITY; D[2 * nelim + 3] = d22;
      }
      tpp_barrier(tg);
      const int nc = n - nelim - 2;
      for (int c = 0; c < nc; ++c) {
         const int cc = nelim + 2 + c;
         const double wc1 = __ldcg(w1 + cc), wc2 = __ldcg(w2 + cc);
         double* ac = Lf + (size_t)cc * ldl;
         for (int r = cc + t0; r < m; r += nt) ac[r] = __ldcg(ac + r) - (__ldcg(a1 + r) * wc1 + __ldcg(a2 + r) * wc2);
      }
      tpp_barrier(tg);
      nelim += 2;
   };
   while (nelim < n) {
      {
         double v[1] = {scan(nelim, nelim, -1)};      // check_col_small(nelim)
         tpp_allmax<1>(tg, v, red);
         if (v[0] < small) { zero_pivot(); continue; }
      }
      bool done = false;
      for (int base = nelim + 1; base < n && !done; base += G) {
         // ---- this CTA's candidate: the whole trial with CTA-local reductions ----
         const int p = base + tg.g;
         int verdict = 0, t = 0;      // 0 rejected, 1 zero column, 2 2x2 pivot (t, p), 3 1x1 pivot p
         double d11 = 0, d21 = 0, d22 = 0;
         if (p < n) {
            double bv = -1.0;
            int bi = p;
            for (int c = nelim + threadIdx.x; c < p; c += TPP_THREADS) {
               const double x = fabs(__ldcg(Lf + (size_t)c * ldl + p));
               if (x > bv) { bv = x; bi = c; }
            }
            double v[2] = {scan1(p, p, -1), bv};
            cta_max2(v, red);
            if (v[0] < small) {
               verdict = 1;
            } else {
               if (threadIdx.x == 0) s_idx = n;
               __syncthreads();
               if (bv == v[1]) atomicMin(&s_idx, bi);
               __syncthreads();
               t = s_idx;
               double mm[2] = {scan1(t, t + 1, p), scan1(p, p + 1, t)};      // maxt, maxp (tpp_rc_max)
               cta_max2(mm, red);
               const double maxt = mm[0];
               double maxp = mm[1];
               // test_2x2 (ldlt_tpp.cxx:89-119)
               const double a11 = __ldcg(Lf + (size_t)t * ldl + t), a21 = __ldcg(Lf + (size_t)t * ldl + p),
                            a22 = __ldcg(Lf + (size_t)p * ldl + p);
               const double maxpiv = fmax(fabs(a11), fmax(fabs(a21), fabs(a22)));
               if (maxpiv >= small) {
                  const double detscale = 1 / maxpiv;
                  const double detpiv0 = (a11 * detscale) * a22;
                  const double detpiv1 = (a21 * detscale) * a21;
                  const double detpiv = detpiv0 - detpiv1;
                  if (!(fabs(detpiv) < fmax(small, fmax(fabs(detpiv0 / 2), fabs(detpiv1 / 2))))) {
                     d11 = (a22 * detscale) / detpiv;
                     d21 = (-a21 * detscale) / detpiv;
                     d22 = (a11 * detscale) / detpiv;
                     if (fmax(maxp, maxt) < small) verdict = 2;
                     else {
                        const double x1 = fabs(d11) * maxt + fabs(d21) * maxp;
                        const double x2 = fabs(d21) * maxt + fabs(d22) * maxp;
                        if (u * fmax(x1, x2) < 1.0) verdict = 2;
                     }
                  }
               }
               if (verdict == 0) {
                  maxp = fmax(maxp, fabs(a21));
                  if (fabs(a22) >= u * maxp) verdict = 3;
               }
            }
         }
         // ---- publish, meet, take the smallest candidate that can be pivoted ----
         const int par = tg.round & 1;
         ++tg.round;
         if (threadIdx.x == 0) {
            double* o = tg.sc->part[par][tg.g];
            o[0] = (double)verdict; o[1] = (double)t; o[2] = d11; o[3] = d21; o[4] = d22;
         }
         tpp_barrier(tg);
         int wg = -1;
         for (int i = 0; i < G && base + i < n; ++i)
            if (__ldcg(&tg.sc->part[par][i][0]) != 0.0) { wg = i; break; }
         if (wg < 0) continue;
         const int wp = base + wg;
         const int wv = (int)__ldcg(&tg.sc->part[par][wg][0]);
         const int wt = (int)__ldcg(&tg.sc->part[par][wg][1]);
         const double e11 = __ldcg(&tg.sc->part[par][wg][2]), e21 = __ldcg(&tg.sc->part[par][wg][3]),
                      e22 = __ldcg(&tg.sc->part[par][wg][4]);
         if (wv == 1) {
            tpp_sym_swap(tg, Lf, Wf, perm, ldl, m, n, nelim, nelim, wp, nelim);
            zero_pivot();
         } else if (wv == 2) {
            tpp_sym_swap(tg, Lf, Wf, perm, ldl, m, n, nelim, nelim, wt, nelim);
            tpp_sym_swap(tg, Lf, Wf, perm, ldl, m, n, nelim, nelim + 1, wp, nelim);
            elim_2x2(e11, e21, e22);
         } else {
            tpp_sym_swap(tg, Lf, Wf, perm, ldl, m, n, nelim, nelim, wp, nelim);
            elim_1x1();
         }
         done = true;
      }
      if (!done) {
         // last resort: 1x1 on column nelim
         double v[1] = {scan(nelim, nelim + 1, -1)};
         tpp_allmax<1>(tg, v, red);
         const double dd = __ldcg(Lf + (size_t)nelim * ldl + nelim);
         if (fabs(dd) >= u * v[0]) {
            elim_1x1();
         } else {
            break;       // out of pivots: the rest is delayed
         }
      }
   }
   tpp_barrier(tg);
   if (threadIdx.x == 0) {
      if (tg.g == 0) {
         st.nelim1 = nelim_in;
         st.nelim = nelim;
         st.p0 = nelim;
      }
      // the barrier counter may only be cleared once every CTA has LEFT the last barrier (a CTA
      // still spinning on it would never see the target again): the last one out clears it
      if (G > 1) {
         __threadfence();
         if (atomicAdd(&st.done, 1) == G - 1) {
            st.arrive = 0;
            st.done = 0;
         }
      }
   }
}

// ---------------------------------------------------------------------------
// k_front_stats: inertia and pivot statistics of a level (src/NumericTree.hxx:150-179;
// factor_failed.hxx:64,118-127).  One thread per front.
// stats: [0] num_delay [1] num_neg [2] num_two [3] num_zero [4] not_first_pass [5] not_second_pass
// ---------------------------------------------------------------------------
// One warp per front, lanes stride the columns: a column is the second of a 2x2 pivot iff its
// first D entry is Inf (the reference's storage convention), so every column classifies itself.
static __global__ void __launch_bounds__(128) k_front_stats(DevTree T, const int* __restrict__ fronts, int cnt,
                                                            int* __restrict__ stats, int* __restrict__ nelim_out,
                                                            const int* __restrict__ slot) {
   const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   const int lane = threadIdx.x & 31;
   if (i >= cnt) return;
   const int f = fronts[i];
   const FrontState st = T.state[f];
   const int n = T.n[f];
   const int nelim = st.nelim;
   const double* d = T.D + T.doff[f];
   int neg = 0, two = 0, zero = 0;
   for (int j = lane; j < nelim; j += 32) {
      const double a11 = d[2 * j], a21 = d[2 * j + 1];
      if (!isfinite(a11)) continue;                          // second column of a 2x2
      if (j + 1 == nelim || isfinite(d[2 * j + 2])) {
         if (a11 == 0.0) ++zero;
         if (a11 < 0.0) ++neg;
      } else {
         const double a22 = d[2 * j + 3];
         ++two;
         const double det = a11 * a22 - a21 * a21;
         const double trace = a11 + a22;
         if (det < 0) ++neg;
         else if (trace < 0) neg += 2;
      }
   }
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      neg += __shfl_xor_sync(0xffffffffu, neg, o);
      two += __shfl_xor_sync(0xffffffffu, two, o);
      zero += __shfl_xor_sync(0xffffffffu, zero, o);
   }
   if (lane != 0) return;
   nelim_out[slot[i]] = nelim;      // the front's position in the level's global list
   if (n - nelim) atomicAdd(&stats[0], n - nelim);
   if (neg) atomicAdd(&stats[1], neg);
   if (two) atomicAdd(&stats[2], two);
   if (zero) atomicAdd(&stats[3], zero);
   if (n - st.nelim1) atomicAdd(&stats[4], n - st.nelim1);
   if (st.nelim1 < n && n - nelim) atomicAdd(&stats[5], n - nelim);
}

// ---------------------------------------------------------------------------
// Assembly for fronts of the indefinite path
// ---------------------------------------------------------------------------
// perm[0:ncol] = rlist[0:ncol]; state reset.  One CTA per front.
static __global__ void __launch_bounds__(256) k_init_front(DevTree T, const int* __restrict__ fronts,
                                                           const int* __restrict__ rlist, const long* __restrict__ rptr) {
   const int f = fronts[blockIdx.x];
   const int ncol0 = T.ncol0[f];
   int* perm = T.perm + T.permoff[f];
   const int* rl = rlist + (rptr[f] - 1);
   for (int i = threadIdx.x; i < ncol0; i += blockDim.x) perm[i] = rl[i];
   if (threadIdx.x == 0) {
      FrontState s;
      s.p0 = 0; s.na = T.n[f]; s.npass = IB; s.kbeg = 0; s.klen = 0; s.wb = 0; s.nelim1 = 0; s.nelim = 0;
      s.obeg = 0; s.oend = 0; s.pa = 0; s.arrive = 0; s.done = 0; s.pad = 0;
      s.cb[0] = s.cb[1] = 0; s.ce[0] = s.ce[1] = 0; s.co[0] = s.co[1] = 0; s.cf[0] = s.cf[1] = 0;
      T.state[f] = s;
   }
}

// A -> front scatter for the fronts of one level (init_a_block, src/kernels/assemble.hxx:162-214).
// grid (fronts, y): the y CTAs of a front stride its slice of the (src,dest) map.
static __global__ void __launch_bounds__(256) k_scatter_a_fronts(DevTree T, const int* __restrict__ fronts,
                                                                 const long* __restrict__ nptr,
                                                                 const long* __restrict__ nlist,
                                                                 const int* __restrict__ nrow0,
                                                                 const double* __restrict__ aval,
                                                                 const double* __restrict__ scaling,
                                                                 const int* __restrict__ rlist,
                                                                 const long* __restrict__ rptr) {
   const int f = fronts[blockIdx.x];
   const int nrow = nrow0[f], ncol0 = T.ncol0[f];
   const int ndelay = T.n[f] - ncol0;
   const int ldl = T.ldl[f];
   double* Lf = T.L + T.loff[f];
   const int* rl = rlist + (rptr[f] - 1);
   for (long e = nptr[f] - 1 + blockIdx.y * (long)blockDim.x + threadIdx.x; e < nptr[f + 1] - 1;
        e += (long)gridDim.y * blockDim.x) {
      const long src = nlist[2 * e] - 1;
      const long dest = nlist[2 * e + 1] - 1;
      const int c = (int)(dest / nrow);
      int r = (int)(dest - (long)c * nrow);
      double v = aval[src];
      if (scaling) v *= scaling[rl[r] - 1] * scaling[rl[c] - 1];
      if (r >= ncol0) r += ndelay;
      Lf[(size_t)c * ldl + r] = v;
   }
}

// Delayed columns of a child -> parent (assemble_delays, src/kernels/assemble.hxx:925-961).
// work item: (child, first delay column in the parent).  One CTA per child.
static __global__ void __launch_bounds__(256) k_assemble_delays(DevTree T, const int2* __restrict__ work) {
   const int2 w = work[blockIdx.x];
   const int c = w.x, dcol = w.y;
   const int p = T.parent[c];
   const int cn = T.n[c], cm_ = T.m[c], cldl = T.ldl[c];
   const int ne = T.state[c].nelim;
   const int nd = cn - ne;
   if (nd <= 0) return;
   const int k = cm_ - cn;
   const double* CL = T.L + T.loff[c];
   const int* cperm = T.perm + T.permoff[c];
   const int* cm = T.cmap + T.cmapoff[c];
   const int pncol0 = T.ncol0[p], pnd = T.n[p] - pncol0, pldl = T.ldl[p];
   double* PL = T.L + T.loff[p];
   int* pperm = T.perm + T.permoff[p];
   for (int j = 0; j < nd; ++j) {
      const double* src = CL + (size_t)(ne + j) * cldl;
      double* dst = PL + (size_t)(dcol + j) * pldl;
      // failed square part (lower)
      for (int i = j + threadIdx.x; i < nd; i += blockDim.x) dst[dcol + i] = src[ne + i];
      // rows of the child's generated element
      for (int i = threadIdx.x; i < k; i += blockDim.x) {
         const int pr = cm[i];
         const double v = src[cn + i];
         if (pr < pncol0) PL[(size_t)pr * pldl + dcol + j] = v;     // above the diagonal: store transposed
         else dst[pr + pnd] = v;
      }
   }
   for (int j = threadIdx.x; j < nd; j += blockDim.x) pperm[dcol + j] = cperm[ne + j];
}

// Extend-add (delay aware variant of k_assemble): parent-local row of child contribution
// row i is cmap[i] (+ ndelay_in of the parent if it is not a fully-summed row).
static __global__ void __launch_bounds__(256) k_assemble_indef(DevTree T, const int2* __restrict__ work, int part) {
   const int2 w = work[blockIdx.x];
   const int c = w.x;
   const int p = T.parent[c];
   const int k = T.m[c] - T.n[c];
   const int* cm = T.cmap + T.cmapoff[c];
   {
      const int je = min(k, w.y + 32);
      if (part == 0 ? (cm[w.y] >= T.ncol0[p]) : (cm[je - 1] < T.ncol0[p])) return;
      if (part == 1 && T.fchild && (T.fchild[2 * p] == c || T.fchild[2 * p + 1] == c)) return;
   }
   const double* src = T.C + T.coff[c];
   const int ldcc = T.ldc[c];
   const int pncol0 = T.ncol0[p], pnd = T.n[p] - pncol0, pldl = T.ldl[p], pldc = T.ldc[p];
   double* PL = T.L + T.loff[p];
   double* PC = T.C + T.coff[p];
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int jend = min(k, w.y + 32);
   for (int j = w.y + warp; j < jend; j += 8) {
      const int rj = cm[j];
      if ((rj < pncol0) != (part == 0)) continue;
      const double* s = src + (size_t)j * ldcc;
      if (rj < pncol0) {
         double* dcol = PL + (size_t)rj * pldl;
         for (int i = j + lane; i < k; i += 32) {
            const int ri = cm[i];
            dcol[ri < pncol0 ? ri : ri + pnd] += s[i];
         }
      } else {
         double* dcol = PC + (size_t)(rj - pncol0) * pldc - pncol0;
         for (int i = j + lane; i < k; i += 32) dcol[cm[i]] += s[i];
      }
   }
}

}  // namespace sylver_b200
