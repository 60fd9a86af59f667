// Dense single-front drivers and micro-benchmarks (helpers group (3) of
// include/sylver_b200.h).  The dense drivers wrap a one-node assembly tree
// around the caller's panel so they exercise exactly the production kernels
// (reference harness: tests/testing_factor_node_posdef.hxx:33-316,
// tests/testing_factor_node_indef.hxx:44-460).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "engine.hpp"
#include "kernels.cuh"

using namespace sylver_b200;

namespace sylver_b200 {
void symbolic_tree_forget(const SymbolicTree* st);
}

// ---------------------------------------------------------------------------
// micro-benchmarks
// ---------------------------------------------------------------------------
// kind 0: register-resident DMMA.8x8x4 issue peak: every warp keeps 16
// independent accumulator pairs in flight.
__global__ void __launch_bounds__(256) k_dmma_peak(double* out, int iters) {
   double c[16][2];
#pragma unroll
   for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
   double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// plain FP64 FMA peak for comparison (kind 2)
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters) {
   double c[16];
#pragma unroll
   for (int i = 0; i < 16; ++i) c[i] = i;
   double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < 16; ++i) s += c[i];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_copy(const double4* __restrict__ a, double4* __restrict__ b, size_t n4) {
   for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
      b[i] = a[i];
}

extern "C" double sylver_b200_bench_copy(long nbytes, int iters) {
   if (device_count() == 0) return -1.0;
   double *a = nullptr, *b = nullptr;
   size_t n4 = (size_t)nbytes / sizeof(double4);
   if (cudaMalloc(&a, n4 * sizeof(double4)) != cudaSuccess || cudaMalloc(&b, n4 * sizeof(double4)) != cudaSuccess) return -1.0;
   cudaMemset(a, 0, n4 * sizeof(double4));
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   k_copy<<<148 * 8, 512>>>((double4*)a, (double4*)b, n4);
   float best = 1e30f;
   for (int i = 0; i < iters; ++i) {
      cudaEventRecord(e0);
      k_copy<<<148 * 8, 512>>>((double4*)a, (double4*)b, n4);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
   }
   cudaFree(a); cudaFree(b);
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   if (cudaGetLastError() != cudaSuccess) return -1.0;
   return 2.0 * n4 * sizeof(double4) / (best * 1e-3) / 1e9;
}

// kind 1: the production tile kernel on one n x n front with k fully-summed
// columns = contribution update (mode 1) of an (n+k) x k panel; returns TFLOP/s
// counting the useful lower-triangle flops n*(n+1)*k.
extern "C" double sylver_b200_bench_dmma(int kind, int n, int k, int iters) {
   if (device_count() == 0) return -1.0;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   double result = -1.0;
   if (kind == 0 || kind == 2) {
      double* out = nullptr;
      const int blocks = 148 * 8;
      cudaMalloc(&out, (size_t)blocks * 256 * sizeof(double));
      const int inner = 4096;
      float best = 1e30f;
      for (int i = 0; i < iters + 1; ++i) {
         cudaEventRecord(e0);
         if (kind == 0) k_dmma_peak<<<blocks, 256>>>(out, inner);
         else k_dfma_peak<<<blocks, 256>>>(out, inner);
         cudaEventRecord(e1);
         cudaEventSynchronize(e1);
         float ms; cudaEventElapsedTime(&ms, e0, e1);
         if (i > 0) best = std::min(best, ms);
      }
      const double flops = (kind == 0) ? (double)blocks * 8 * inner * 16 * 512.0
                                       : (double)blocks * 256 * inner * 16 * 2.0;
      result = flops / (best * 1e-3) / 1e12;
      cudaFree(out);
   } else if (kind == 1) {
      // one front: m = n + k rows, k columns, contribution n x n
      const int m = n + k;
      std::vector<int> hm{m}, hn{k}, hldl{(m + 3) / 4 * 4}, hldc{(n + 3) / 4 * 4}, hpar{1}, hnch{0, 0};
      std::vector<long> hloff{0}, hcoff{0}, hcmo{0, 0};
      int *dm, *dn, *dldl, *dldc, *dpar, *dnch, *dfr, *dpre;
      long *dloff, *dcoff, *dcmo;
      double *L, *C;
      cudaMalloc(&dm, 4); cudaMalloc(&dn, 4); cudaMalloc(&dldl, 4); cudaMalloc(&dldc, 4); cudaMalloc(&dpar, 4);
      cudaMalloc(&dnch, 8); cudaMalloc(&dfr, 4); cudaMalloc(&dloff, 8); cudaMalloc(&dcoff, 8); cudaMalloc(&dcmo, 16);
      cudaMemcpy(dm, hm.data(), 4, cudaMemcpyHostToDevice); cudaMemcpy(dn, hn.data(), 4, cudaMemcpyHostToDevice);
      cudaMemcpy(dldl, hldl.data(), 4, cudaMemcpyHostToDevice); cudaMemcpy(dldc, hldc.data(), 4, cudaMemcpyHostToDevice);
      cudaMemcpy(dpar, hpar.data(), 4, cudaMemcpyHostToDevice); cudaMemcpy(dnch, hnch.data(), 8, cudaMemcpyHostToDevice);
      cudaMemcpy(dloff, hloff.data(), 8, cudaMemcpyHostToDevice); cudaMemcpy(dcoff, hcoff.data(), 8, cudaMemcpyHostToDevice);
      cudaMemcpy(dcmo, hcmo.data(), 16, cudaMemcpyHostToDevice);
      int zero = 0; cudaMemcpy(dfr, &zero, 4, cudaMemcpyHostToDevice);
      const int base = k & ~1;
      const int TR = (m - base + GT_BM - 1) / GT_BM;
      int tiles = 0;
      for (int tj = 0; tj < TR; ++tj) tiles += TR - tj;
      int pre[2] = {0, tiles};
      cudaMalloc(&dpre, 8); cudaMemcpy(dpre, pre, 8, cudaMemcpyHostToDevice);
      cudaMalloc(&L, (size_t)hldl[0] * k * 8); cudaMalloc(&C, (size_t)hldc[0] * n * 8);
      cudaMemset(L, 0, (size_t)hldl[0] * k * 8);
      DevTree T{dm, dn, dldl, dldc, dloff, dcoff, dcmo, dpar, dnch, nullptr, L, C};
      TileBatch b{dfr, dpre, 1};
      cudaFuncSetAttribute(k_gemm_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
      cudaFuncSetAttribute(k_gemm_batched, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (getenv("SYLVER_B200_VERBOSE")) {
         int nbk = 0;
         cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbk, k_gemm_batched, GT_THREADS, GT_SMEM_BYTES);
         fprintf(stderr, "sylver_b200: k_gemm_batched resident CTAs per SM: %d\n", nbk);
      }
      float best = 1e30f;
      for (int i = 0; i < iters + 1; ++i) {
         cudaEventRecord(e0);
         k_gemm_batched<<<gemm_grid(1, tiles), GT_THREADS, GT_SMEM_BYTES>>>(T, b, 1, 0, 128, nullptr, 0, 0, 1);
         cudaEventRecord(e1);
         cudaEventSynchronize(e1);
         float ms; cudaEventElapsedTime(&ms, e0, e1);
         if (i > 0) best = std::min(best, ms);
      }
      result = (double)n * (n + 1) * k / (best * 1e-3) / 1e12;
      cudaFree(dm); cudaFree(dn); cudaFree(dldl); cudaFree(dldc); cudaFree(dpar); cudaFree(dnch); cudaFree(dfr);
      cudaFree(dpre); cudaFree(dloff); cudaFree(dcoff); cudaFree(dcmo); cudaFree(L); cudaFree(C);
   }
   else if (kind == 3 || kind == 4) {
      // latency of the diagonal-block kernels on ONE 128 x 128 SPD block (microseconds per launch):
      // kind 3 = k_potrf_inv_reg (column at a time), kind 4 = k_potrf_inv_blk (blocked)
      const int m = 128, ld = 128;
      std::vector<double> h((size_t)ld * m, 0.0);
      for (int c = 0; c < m; ++c)
         for (int r = c; r < m; ++r) h[(size_t)c * ld + r] = (r == c) ? 2.0 * m : 1.0 / (1.0 + r - c);
      std::vector<int> hm{m}, hn{m}, hl{ld};
      std::vector<long> ho{0};
      int *dm, *dn, *dl, *dfr, *dfail;
      long* dlo;
      double *L, *L0, *Wd;
      cudaMalloc(&dm, 4); cudaMalloc(&dn, 4); cudaMalloc(&dl, 4); cudaMalloc(&dfr, 4); cudaMalloc(&dfail, 16);
      cudaMalloc(&dlo, 8); cudaMalloc(&L, h.size() * 8); cudaMalloc(&L0, h.size() * 8); cudaMalloc(&Wd, (size_t)m * m * 8);
      cudaMemcpy(dm, hm.data(), 4, cudaMemcpyHostToDevice); cudaMemcpy(dn, hn.data(), 4, cudaMemcpyHostToDevice);
      cudaMemcpy(dl, hl.data(), 4, cudaMemcpyHostToDevice); cudaMemcpy(dlo, ho.data(), 8, cudaMemcpyHostToDevice);
      cudaMemset(dfr, 0, 4); cudaMemset(dfail, 0, 16);
      cudaMemcpy(L0, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
      DevTree T{};
      T.m = dm; T.n = dn; T.ldl = dl; T.loff = dlo; T.L = L;
      cudaFuncSetAttribute(k_potrf_inv_blk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PB_SMEM_BYTES);
      float best = 1e30f;
      for (int i = 0; i < iters + 1; ++i) {
         cudaMemcpy(L, L0, h.size() * 8, cudaMemcpyDeviceToDevice);
         cudaDeviceSynchronize();
         cudaEventRecord(e0);
         if (kind == 3) k_potrf_inv_reg<<<1, PR_THREADS>>>(T, dfr, 0, 128, Wd, 128, dfail);
         else k_potrf_inv_blk<<<1, PB_THREADS, PB_SMEM_BYTES>>>(T, dfr, 0, 128, Wd, 128, dfail);
         cudaEventRecord(e1);
         cudaEventSynchronize(e1);
         float ms; cudaEventElapsedTime(&ms, e0, e1);
         if (i > 0) best = std::min(best, ms);
      }
      result = best * 1e3;
      cudaFree(dm); cudaFree(dn); cudaFree(dl); cudaFree(dfr); cudaFree(dfail); cudaFree(dlo); cudaFree(L); cudaFree(L0); cudaFree(Wd);
   }
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   if (cudaGetLastError() != cudaSuccess) return -1.0;
   return result;
}

// ---------------------------------------------------------------------------
// dense single fronts
// ---------------------------------------------------------------------------
static SymbolicTree* one_front_tree(int m, int n, int lda) {
   // system of order m, a single supernode of n columns whose row list is 1..m
   std::vector<int> sptr{1, n + 1}, sparent{2}, rlist(m);
   std::vector<long> rptr{1, (long)m + 1}, nptr(2), nlist;
   for (int i = 0; i < m; ++i) rlist[i] = i + 1;
   nlist.reserve((size_t)2 * ((size_t)n * m));
   for (int c = 0; c < n; ++c)
      for (int r = c; r < m; ++r) {
         nlist.push_back((long)c * lda + r + 1);
         nlist.push_back((long)c * m + r + 1);
      }
   nptr[0] = 1;
   nptr[1] = (long)nlist.size() / 2 + 1;
   int flag = 0;
   return symbolic_tree_create(m, 1, sptr.data(), sparent.data(), rptr.data(), rlist.data(), nptr.data(),
                               nlist.data(), &flag);
}

extern "C" int sylver_b200_factor_front_posdef(int m, int n, double* a, int lda, double* contrib, int nb,
                                                float* ms_out) {
   (void)nb;
   if (device_count() == 0) return SYLVER_ERROR_CUDA_UNKNOWN;
   SymbolicTree* st = one_front_tree(m, n, lda);
   if (!st) return SYLVER_ERROR_UNKNOWN;
   sylver_options_c opt{};
   opt.nb = 256; opt.u = 0.01; opt.small = 1e-20; opt.action = true; opt.multiplier = 1.1;
   opt.pivot_method = 2; opt.failed_pivot_method = 1;
   sylver_inform_c stats{};
   NumericTree* nt = numeric_tree_create(true, st, a, nullptr, &opt, &stats);
   int ret = n;
   if (!nt) {
      ret = stats.flag ? stats.flag : SYLVER_ERROR_UNKNOWN;
   } else {
      if (stats.flag < 0) ret = stats.flag;
      double t[4];
      numeric_tree_timings(nt, t);
      if (ms_out) *ms_out = (float)(t[0] * 1e3);
      std::vector<double> l((size_t)m * n);
      int mm, nn;
      if (numeric_tree_get_front(nt, 0, &mm, &nn, l.data(), contrib) != 0) ret = SYLVER_ERROR_CUDA_UNKNOWN;
      for (int c = 0; c < n; ++c)
         for (int r = c; r < m; ++r) a[(size_t)c * lda + r] = l[(size_t)c * m + r];
      numeric_tree_destroy(nt);
   }
   symbolic_tree_forget(st);
   delete st;
   return ret;
}

// Indefinite single front (reference harness tests/testing_factor_node_indef.hxx:44-460 drives
// factor_front_indef the same way): APTP on the n fully-summed columns of an m x m symmetric
// front, TPP on what fails, the rest is reported as delayed (nelim < n).
extern "C" int sylver_b200_factor_front_indef(int m, int n, int* perm, double* a, int lda, double* d, double* contrib,
                                               sylver_options_c const* options, sylver_inform_c* stats,
                                               float* ms_out) {
   if (device_count() == 0) return SYLVER_ERROR_CUDA_UNKNOWN;
   SymbolicTree* st = one_front_tree(m, n, lda);
   if (!st) return SYLVER_ERROR_UNKNOWN;
   sylver_inform_c local{};
   if (!stats) stats = &local;
   NumericTree* nt = numeric_tree_create(false, st, a, nullptr, options, stats);
   int ret;
   if (!nt) {
      ret = stats->flag ? stats->flag : SYLVER_ERROR_UNKNOWN;
   } else {
      // the reported time is that of a second factorization on the same tree: the first one grows
      // the level scratch buffers (cudaMalloc / cudaMallocHost between launches), which lands
      // inside the event-bracketed interval and varies from 0 to 100 ms
      if (ms_out && stats->flag >= 0) numeric_tree_refactor(nt, a, nullptr, options, stats);
      double t[4];
      numeric_tree_timings(nt, t);
      if (ms_out) *ms_out = (float)(t[0] * 1e3);
      std::vector<double> l((size_t)m * n);
      std::vector<int> idx(n);
      int mm, nn, nelim = 0;
      ret = 0;
      if (numeric_tree_get_front(nt, 0, &mm, &nn, l.data(), contrib) != 0 ||
          numeric_tree_get_front_indef(nt, 0, &nelim, d, idx.data()) != 0)
         ret = SYLVER_ERROR_CUDA_UNKNOWN;
      if (ret == 0) {
         for (int c = 0; c < n; ++c)
            for (int r = c; r < m; ++r) a[(size_t)c * lda + r] = l[(size_t)c * m + r];
         if (perm) {
            std::vector<int> in(perm, perm + n);
            for (int i = 0; i < n; ++i) perm[i] = in[idx[i] - 1];
         }
         ret = (stats->flag < 0) ? stats->flag : nelim;
      }
      numeric_tree_destroy(nt);
   }
   symbolic_tree_forget(st);
   delete st;
   return ret;
}
