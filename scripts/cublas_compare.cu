// Same-box comparison the survey sets as the bar: what the reference's GPU tile
// tasks call per tile -- cublasDgemm / cublasDsyrk / cublasDtrsm (src/StarPU/cuda/kernels.hxx:45-182)
// -- timed on this B200 on (a) the reference's own per-tile shapes (nb = 256) and (b) the whole
// panel updates the B200 engine issues as one launch (K = 256 trailing update, K = 5000
// contribution update of a 10000-row block).  Build + run (GPU box):
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a scripts/cublas_compare.cu -lcublas -o /tmp/cublas_compare && /tmp/cublas_compare
// Not product code: libsylver_b200.so never links cuBLAS.
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

static double time_ms(cudaStream_t s, int iters, const auto& f) {
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   f();                       // warm-up
   cudaStreamSynchronize(s);
   float best = 1e30f;
   for (int i = 0; i < iters; ++i) {
      cudaEventRecord(e0, s);
      f();
      cudaEventRecord(e1, s);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
   }
   return best;
}

int main() {
   cublasHandle_t h;
   cublasCreate(&h);
   cudaStream_t s;
   cudaStreamCreate(&s);
   cublasSetStream(h, s);
   const int N = 10240, K = 5120;
   double *A, *B, *Cm;
   cudaMalloc(&A, (size_t)N * K * 8); cudaMalloc(&B, (size_t)N * K * 8); cudaMalloc(&Cm, (size_t)N * N * 8);
   cudaMemset(A, 0, (size_t)N * K * 8); cudaMemset(B, 0, (size_t)N * K * 8); cudaMemset(Cm, 0, (size_t)N * N * 8);
   // a well conditioned triangular block for trsm
   {
      std::vector<double> t((size_t)256 * 256, 0.0);
      for (int i = 0; i < 256; ++i) t[(size_t)i * 256 + i] = 4.0;
      cudaMemcpy(B, t.data(), t.size() * 8, cudaMemcpyHostToDevice);
   }
   const double one = 1.0, mone = -1.0;
   printf("{");
   // (a) the reference's per-tile calls, nb = 256 (src/StarPU/cuda/kernels.hxx): one tile at a time
   {
      const int nb = 256;
      double ms = time_ms(s, 20, [&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, nb, nb, nb, &mone, A, N, A + nb, N, &one, Cm, N); });
      printf("\"dgemm_tile_256\": {\"ms\": %.4f, \"tflops\": %.2f}, ", ms, 2.0 * nb * nb * nb / ms / 1e9);
      ms = time_ms(s, 20, [&] { cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, nb, nb, &mone, A, N, &one, Cm, N); });
      printf("\"dsyrk_tile_256\": {\"ms\": %.4f, \"tflops\": %.2f}, ", ms, 1.0 * nb * nb * nb / ms / 1e9);
      ms = time_ms(s, 20, [&] { cublasDtrsm(h, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, nb, nb, &one, B, 256, A, N); });
      printf("\"dtrsm_tile_256\": {\"ms\": %.4f, \"tflops\": %.2f}, ", ms, 1.0 * nb * nb * nb / ms / 1e9);
   }
   // (b) whole-panel shapes (what one launch of the B200 engine covers)
   for (int k : {256, 5000}) {
      const int n = 10000;
      double ms = time_ms(s, 5, [&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, k, &mone, A, N, B, N, &one, Cm, N); });
      printf("\"dgemm_NT_%dx%dx%d\": {\"ms\": %.3f, \"tflops\": %.2f}, ", n, n, k, ms, 2.0 * n * n * k / ms / 1e9);
      ms = time_ms(s, 5, [&] { cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, k, &mone, A, N, &one, Cm, N); });
      printf("\"dsyrk_L_%dx%d\": {\"ms\": %.3f, \"tflops_useful\": %.2f}, ", n, k, ms, 1.0 * n * (n + 1) * k / ms / 1e9);
   }
   {
      const int m = 10000, nb = 128;
      double ms = time_ms(s, 10, [&] { cublasDtrsm(h, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, m, nb, &one, B, 256, A, N); });
      printf("\"dtrsm_R_L_T_%dx%d\": {\"ms\": %.4f, \"tflops\": %.2f}", m, nb, ms, 1.0 * m * nb * nb / ms / 1e9);
   }
   printf("}\n");
   return 0;
}
