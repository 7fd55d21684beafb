// Bring-up probe: tcgen05.mma kind::f16 with MN-major A and B straight from row-major fp16 tiles
// (the layout of a spilled K panel [rows x centres] and of W [rows x T]): D[m, n] = sum_k A[k][m] B[k][n].
// Verifies the shared-memory descriptor (LBO / SBO roles) against a CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/mn_probe tools/mn_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "../online-detection_b200/csrc/odf_ptx.cuh"

using namespace odf;

constexpr int R = 64;          // rows (K) per stage
constexpr int MT = 128;        // centres per tile (M)
constexpr int NT = 64;         // rhs columns (N)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(const void* base, int rows, int cols, int ld, bool bf16 = false) {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)R};
  cuuint32_t es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)sym)(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_s = smem;                       // 2 chunks of [R x 128 B] (64 fp16 along M each)
  uint8_t* b_s = smem + 4 * R * 128;         // 1 chunk
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * R * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(bars), 1); mbar_init(smem_u32(bars + 1), 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(smem_u32(slot), 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(smem_u32(bars), 3 * R * 128);
    for (int j = 0; j < 2; ++j) tma_load_2d(smem_u32(a_s + j * R * 128), &tmA, smem_u32(bars), j * 64, 0);
    for (int j = 0; j < 1; ++j) tma_load_2d(smem_u32(b_s + j * R * 128), &tmB, smem_u32(bars), j * 64, 0);
    mbar_wait(smem_u32(bars), 0);
    tc_fence_after();
    // kind::f16 (fp16 x fp16), fp32 accumulate, A and B MN-major (bits 15 / 16), N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((NT >> 3) << 17) | ((MT >> 4) << 24);
    const uint32_t chunk = R * 128, grp = 1024;
    for (int kk = 0; kk < R / 16; ++kk) {      // K = 16 per fp16 MMA = two 8-row k-groups
      uint64_t da, db;
      if (variant == 0) { da = make_desc(smem_u32(a_s) + kk * 2 * grp, chunk, grp); db = make_desc(smem_u32(b_s) + kk * 2 * grp, chunk, grp); }
      else              { da = make_desc(smem_u32(a_s) + kk * 2 * grp, grp, chunk); db = make_desc(smem_u32(b_s) + kk * 2 * grp, grp, chunk); }
      mma_f16_ss(tmem, da, db, idesc, kk > 0 ? 1u : 0u);
    }
    tc_commit(smem_u32(bars + 1));
    mbar_wait(smem_u32(bars + 1), 0);
  }
  __syncthreads();
  tc_fence_after();
  // epilogue: thread = TMEM lane = m
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
  for (int c0 = 0; c0 < NT; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + lane_off + c0, r);
    tc_wait_ld();
    for (int c = 0; c < 32; ++c) out[threadIdx.x * NT + c0 + c] = __uint_as_float(r[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<float> A(R * MT), B(R * NT); std::vector<__half> Ah(R * MT); std::vector<__half> Bh(R * NT);
  for (int k = 0; k < R; ++k) {
    for (int m = 0; m < MT; ++m) A[k * MT + m] = float((k * 7 + m * 3) % 11 - 5);
    for (int n = 0; n < NT; ++n) B[k * NT + n] = float((k * 5 + n * 13) % 9 - 4);
  }
  for (size_t i = 0; i < A.size(); ++i) Ah[i] = __float2half(A[i]);
  for (size_t i = 0; i < B.size(); ++i) Bh[i] = __float2half(B[i]);
  __half *dA; __half* dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, MT * NT * 4);
  cudaMemcpy(dA, Ah.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bh.data(), B.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tA = make_map(dA, R, MT, MT), tB = make_map(dB, R, NT, NT);
  const int smem = 6 * R * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dO, 0, MT * NT * 4);
    probe<<<1, 128, smem>>>(tA, tB, dO, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
    std::vector<float> O(MT * NT);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < MT; ++m)
      for (int n = 0; n < NT; ++n) {
        double ref = 0;
        if (variant == 2) { for (int k = 0; k < 32; ++k) ref += double(A[(m % 64) * MT + (m / 64) * 32 + k]) * B[n * NT + k]; }
        else for (int k = 0; k < R; ++k) ref += double(A[k * MT + m]) * B[k * NT + n];
        const double err = fabs(ref - O[m * NT + n]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3) ++bad;
      }
    for (int m = 0; m < 3; ++m) { printf("  m=%d:", m); for (int n = 0; n < 6; ++n) { double ref = 0; for (int k = 0; k < R; ++k) ref += double(A[k * MT + m]) * B[k * NT + n]; printf(" %g/%g", O[m * NT + n], ref);} printf("\n"); }
    printf("variant %d (%s): max abs err %.3f, mismatching entries %d / %d\n", variant,
           variant == 0 ? "LBO = MN-chunk stride, SBO = 8-row k-group stride" : "LBO = k-group stride, SBO = MN-chunk stride",
           maxerr, bad, MT * NT);
  }
  return 0;
}
