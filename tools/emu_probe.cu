// Probe: cuBLAS FP32 emulation on bf16 tensor cores (BF16x9) vs SIMT sgemm / ssyrk / strsm and cuSOLVER potrf.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++20 -o tools/bin/emu_probe tools/emu_probe.cu -lcublas -lcusolver
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <cuda_runtime.h>

#define CK(x) do { auto e_ = (x); if (e_ != 0) { printf("FAIL %s -> %d (line %d)\n", #x, (int)e_, __LINE__); } } while (0)

static float timeit(cudaStream_t st, int reps, auto fn) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  fn(); cudaStreamSynchronize(st);
  cudaEventRecord(a, st);
  for (int i = 0; i < reps; ++i) fn();
  cudaEventRecord(b, st); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

__global__ void fill(float* A, long n, unsigned seed) {
  long i = blockIdx.x * 256L + threadIdx.x;
  if (i < n) { unsigned x = (unsigned)i * 2654435761u + seed; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; A[i] = (x & 0xffff) / 65536.f - 0.5f; }
}
__global__ void add_diag(float* A, int m, float v) { int i = blockIdx.x * 256 + threadIdx.x; if (i < m) A[(long)i * m + i] += v; }

int main(int argc, char** argv) {
  int m = argc > 1 ? atoi(argv[1]) : 10000;
  cublasHandle_t h; cublasCreate(&h);
  cusolverDnHandle_t s; cusolverDnCreate(&s);
  cudaStream_t st; cudaStreamCreate(&st); cublasSetStream(h, st); cusolverDnSetStream(s, st);
  float *A, *B, *C, *C2; long n = (long)m * m;
  cudaMalloc(&A, n * 4); cudaMalloc(&B, n * 4); cudaMalloc(&C, n * 4); cudaMalloc(&C2, n * 4);
  fill<<<(n + 255) / 256, 256, 0, st>>>(A, n, 1); fill<<<(n + 255) / 256, 256, 0, st>>>(B, n, 2);
  const float one = 1.f, zero = 0.f;
  double fl = 2.0 * m * (double)m * m;
  float t0 = timeit(st, 2, [&] { CK(cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, m, m, m, &one, A, m, B, m, &zero, C, m)); });
  printf("m=%d sgemm default: %.2f ms  %.1f TFLOP/s\n", m, t0, fl / t0 / 1e9);
  float t1 = timeit(st, 2, [&] { CK(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_T, m, m, m, &one, A, CUDA_R_32F, m, B, CUDA_R_32F, m, &zero, C2, CUDA_R_32F, m,
                                              CUBLAS_COMPUTE_32F_EMULATED_16BFX9, CUBLAS_GEMM_DEFAULT)); });
  printf("m=%d gemmEx EMULATED_16BFX9: %.2f ms  %.1f TFLOP/s\n", m, t1, fl / t1 / 1e9);
  // accuracy of both against fp64 on a few entries
  {
    std::vector<float> ha(n), hb(n); std::vector<float> c1(m), c2(m);
    cudaMemcpy(ha.data(), A, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), B, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(c1.data(), C, m * 4, cudaMemcpyDeviceToHost); cudaMemcpy(c2.data(), C2, m * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0, sc = 0;
    for (int i = 0; i < 64; ++i) {   // column 0 of C (column-major): C[i,0] = sum_k A[i,k] B[0,k]
      double r = 0, ab = 0; for (int k = 0; k < m; ++k) { r += (double)ha[(long)k * m + i] * hb[(long)k * m]; ab += fabs((double)ha[(long)k * m + i] * hb[(long)k * m]); }
      e1 = fmax(e1, fabs(r - c1[i])); e2 = fmax(e2, fabs(r - c2[i])); sc = fmax(sc, ab);
    }
    printf("   max abs err / sum|ab|: simt %.2e  emulated %.2e\n", e1 / sc, e2 / sc);
  }
  CK(cublasSetMathMode(h, CUBLAS_FP32_EMULATED_BF16X9_MATH));
  CK(cublasSetEmulationStrategy(h, CUBLAS_EMULATION_STRATEGY_EAGER));
  float t2 = timeit(st, 2, [&] { CK(cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, m, m, m, &one, A, m, B, m, &zero, C, m)); });
  printf("m=%d sgemm + BF16X9 math mode (eager): %.2f ms  %.1f TFLOP/s\n", m, t2, fl / t2 / 1e9);
  float t3 = timeit(st, 2, [&] { CK(cublasSsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, m, m, &one, A, m, &zero, C, m)); });
  printf("m=%d ssyrk + math mode: %.2f ms  %.1f TFLOP/s (m^3)\n", m, t3, fl / 2 / t3 / 1e9);
  float t4 = timeit(st, 1, [&] { CK(cublasStrsm(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, m, m, &one, A, m, C2, m)); });
  printf("m=%d strsm (m rhs) + math mode: %.2f ms  %.1f TFLOP/s (m^3)\n", m, t4, fl / 2 / t4 / 1e9);
  CK(cublasSetMathMode(h, CUBLAS_DEFAULT_MATH));
  float t5 = timeit(st, 1, [&] { CK(cublasSsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, m, m, &one, A, m, &zero, C, m)); });
  printf("m=%d ssyrk default: %.2f ms  %.1f TFLOP/s (m^3)\n", m, t5, fl / 2 / t5 / 1e9);
  float t6 = timeit(st, 1, [&] { CK(cublasStrsm(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, m, m, &one, A, m, C2, m)); });
  printf("m=%d strsm default: %.2f ms  %.1f TFLOP/s (m^3)\n", m, t6, fl / 2 / t6 / 1e9);
  // potrf on SPD: C = A^T A + m I
  CK(cublasSsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, m, m, &one, A, m, &zero, C, m));
  add_diag<<<(m + 255) / 256, 256, 0, st>>>(C, m, (float)m);
  int lwork = 0; CK(cusolverDnSpotrf_bufferSize(s, CUBLAS_FILL_MODE_LOWER, m, C, m, &lwork));
  float* work; cudaMalloc(&work, (size_t)lwork * 4); int* info; cudaMalloc(&info, 4);
  cudaMemcpyAsync(C2, C, n * 4, cudaMemcpyDeviceToDevice, st);
  float t7 = timeit(st, 1, [&] { cudaMemcpyAsync(C, C2, n * 4, cudaMemcpyDeviceToDevice, st); CK(cusolverDnSpotrf(s, CUBLAS_FILL_MODE_LOWER, m, C, m, work, lwork, info)); });
  printf("m=%d cusolver spotrf (+copy): %.2f ms  %.1f TFLOP/s (m^3/3)\n", m, t7, fl / 6 / t7 / 1e9);
  cudaDeviceSynchronize();
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
