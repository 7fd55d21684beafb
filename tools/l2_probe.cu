// L2 retention probe (B200): pass 1 streams a buffer of S MB with one CTA per SM (CTA b reads slice b), pass 2 reads it again with
// CTA b reading slice (b + shift) % grid.  Loads bypass L1 (ld.global.cg).  If the L2 is one cache for all SMs, pass 2 runs at L2
// speed whatever the shift; if each die keeps its own copy of what ITS SMs read, a shifted pass 2 misses when the slice was first
// touched from the other die.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/l2_probe tools/l2_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) read_slices(const uint4* __restrict__ buf, size_t slice_vec, int shift, unsigned long long* sink, unsigned* smid_of) {
  extern __shared__ unsigned char pad[];
  const int b = (blockIdx.x + shift) % gridDim.x;
  const uint4* p = buf + static_cast<size_t>(b) * slice_vec;
  unsigned acc = 0;
  for (size_t i = threadIdx.x; i < slice_vec; i += blockDim.x) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
    acc += v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
  if (threadIdx.x == 0 && smid_of) {
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    smid_of[blockIdx.x] = s;
  }
}

int main(int argc, char** argv) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(read_slices, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  unsigned long long* sink;
  cudaMalloc(&sink, 8);
  unsigned* smid;
  cudaMalloc(&smid, sms * 4);
  uint4* flush;
  const size_t flush_bytes = 512ull << 20;
  cudaMalloc(&flush, flush_bytes);
  for (int mb : {8, 16, 32, 48, 64, 96, 120}) {
    const size_t slice_vec = (static_cast<size_t>(mb) << 20) / sms / 16 / 512 * 512;
    const size_t bytes = slice_vec * 16 * sms;
    uint4* buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 1, bytes);
    for (int shift : {0, 1, 2, sms / 4, sms / 2, sms - 1}) {
      float best1 = 1e9f, best2 = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(flush, rep, flush_bytes);                              // evict the buffer
        cudaEvent_t e0, e1, e2;
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
        cudaEventRecord(e0);
        read_slices<<<sms, 512, 200 * 1024>>>(buf, slice_vec, 0, sink, smid);
        cudaEventRecord(e1);
        read_slices<<<sms, 512, 200 * 1024>>>(buf, slice_vec, shift, sink, nullptr);
        cudaEventRecord(e2);
        cudaDeviceSynchronize();
        float t1, t2;
        cudaEventElapsedTime(&t1, e0, e1);
        cudaEventElapsedTime(&t2, e1, e2);
        if (t1 < best1) best1 = t1;
        if (t2 < best2) best2 = t2;
      }
      printf("buffer %3d MB  shift %3d: first touch %.1f us (%.0f GB/s)   second touch %.1f us (%.0f GB/s)\n", mb, shift, best1 * 1e3,
             bytes / best1 * 1e-6, best2 * 1e3, bytes / best2 * 1e-6);
    }
    cudaFree(buf);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
