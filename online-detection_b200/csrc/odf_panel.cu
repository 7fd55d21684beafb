// Second half of a K^T (K v) sweep without recomputing K: the fused tile spills its K tiles
// (fp32, K_hi + K_lo exactly as the first contraction used them) into a transient row PANEL
// P [rows x ldp] and this kernel contracts it with the finished W = K v + w of the same rows:
//     out_partial[split][c][t] = sum_{r in split} P[r][c] * W[r][t]
// It replaces the second call of falkon `GaussianKernel.dmmv`'s inner product K_blk^T w
// (reached from InCoreFalkon.fit, FALKONWrapper_with_centers_selection_incore.py:68).
//
// The contraction has T <= 32 columns, i.e. 16 flop per panel byte: it is bound by streaming the
// panel from HBM once and by the fp32 FMA pipe, not by the tensor cores, so it is a register-tiled
// fp32 kernel (exact fp32 products, no operand split): 128 threads own a [256 centres x T_pad]
// accumulator tile (8 centres x 8 (or 4) right-hand sides per thread, kept as packed f32x2 pairs and
// updated with fma.rn.f32x2 — two FMAs per issue slot on sm_100).  Panel rows arrive as 1 KB
// cp.async.bulk copies (one elected thread issues a whole stage, an mbarrier counts the bytes), three
// stages deep; panel reads are conflict-free 32-byte-per-lane LDS.128, W reads are warp broadcasts.
// Splits over rows write separate slabs that odf_finish_rows reduces in index order (deterministic).
#include <cstdlib>
#include "odf_internal.h"
#include "odf_ptx.cuh"

namespace odf {
namespace {

constexpr int PC = 256;   // centres per CTA
constexpr int PR = 32;    // rows per stage
constexpr int PSTAGES = 3;
constexpr int PANEL_SMEM = PSTAGES * PR * (PC + 32) * 4 + PSTAGES * 8 + 16;   // 110 KB -> 2 CTAs / SM

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void fma2(uint64_t& acc, uint64_t a, uint64_t b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

template <int TP, int PACKED>
__global__ void __launch_bounds__(256, 2)
panel_tmm_kernel(const __grid_constant__ CUtensorMap tmP, const float* __restrict__ W, int64_t n_rows, int M,
                 int64_t rows_per_split, float* __restrict__ out_partial) {
  constexpr int NP = TP / 8;                     // f32x2 pairs of right-hand sides per thread (4 or 2)
  extern __shared__ __align__(128) uint8_t psm[];
  float* Ps = reinterpret_cast<float*>(psm);                                     // [PSTAGES][PR][PC]
  float* Ws = Ps + PSTAGES * PR * PC;                                            // [PSTAGES][PR][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ws + PSTAGES * PR * 32);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // warps 0-3 take the even rows of a stage, warps 4-7 the odd rows (same accumulator tiles, summed at
  // the end): 16 resident warps per SM hide the LDS latency that 8 could not
  const int cg = tid & 31, tg = warp & 3, rg = warp >> 2;   // 32 centre groups of 8; 4 rhs groups of 2*NP
  const int c0 = blockIdx.x * PC;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.y) * rows_per_split;
  const int64_t r_end = min(n_rows, r_begin + rows_per_split);
  const int n_stages = static_cast<int>((r_end - r_begin + PR - 1) / PR);

  if (tid == 0) {
    for (int s = 0; s < PSTAGES; ++s) mbar_init(smem_u32(bars + s), 1);
    fence_barrier_init();
  }
  __syncthreads();

  // one elected lane of warp 0 (warp-uniform branch, see elect_one()): a [PR x PC] TMA box of the panel
  // (rows / columns past the end of the panel read as zeros) + one bulk copy of the W rows
  auto issue = [&](int st) {
    if (st >= n_stages) return;
    const int buf = st % PSTAGES;
    const int64_t r0 = r_begin + static_cast<int64_t>(st) * PR;
    const int rows = static_cast<int>(min(static_cast<int64_t>(PR), r_end - r0));
    const uint32_t bar = smem_u32(bars + buf);
    mbar_arrive_expect_tx(bar, static_cast<uint32_t>(PR * PC * 4 + rows * TP * 4));
    tma_load_2d(smem_u32(Ps + buf * PR * PC), &tmP, bar, c0, static_cast<int>(r0));
    bulk_g2s(smem_u32(Ws + buf * PR * 32), W + r0 * TP, rows * TP * 4, bar);     // W rows are contiguous
  };

  uint64_t acc[8][NP];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NP; ++j) acc[i][j] = 0ull;

  if (warp == 0) {
    if (elect_one())
      for (int s = 0; s < PSTAGES - 1; ++s) issue(s);
    __syncwarp();
  }
  for (int st = 0; st < n_stages; ++st) {
    const int buf = st % PSTAGES;
    if (warp == 0) {                             // into the buffer everyone left at the end of iteration st-1
      if (elect_one()) issue(st + PSTAGES - 1);
      __syncwarp();
    }
    mbar_wait(smem_u32(bars + buf), (st / PSTAGES) & 1);
    const int64_t r0 = r_begin + static_cast<int64_t>(st) * PR;
    const int rows = static_cast<int>(min(static_cast<int64_t>(PR), r_end - r0));
    // lane cg owns centres [4cg, 4cg+4) and [128+4cg, 128+4cg+4): each LDS.128 of a quarter-warp covers
    // 128 contiguous bytes (lanes 32 bytes apart would collide two-way on the banks)
    const float* ps = Ps + buf * PR * PC + cg * 4;
    const float* ws = Ws + buf * PR * 32 + tg * (2 * NP);       // W tile row pitch is TP floats
#pragma unroll 4
    for (int r = rg; r < rows; r += 2) {
      const float4 p0 = *reinterpret_cast<const float4*>(ps + r * PC);
      const float4 p1 = *reinterpret_cast<const float4*>(ps + r * PC + 128);
      uint64_t w[NP];
      if (NP == 4) {
        const ulonglong2 t0 = *reinterpret_cast<const ulonglong2*>(ws + r * TP);
        const ulonglong2 t1 = *reinterpret_cast<const ulonglong2*>(ws + r * TP + 4);
        w[0] = t0.x; w[1] = t0.y; w[2] = t1.x; w[NP - 1] = t1.y;
      } else {
        const ulonglong2 t0 = *reinterpret_cast<const ulonglong2*>(ws + r * TP);
        w[0] = t0.x; w[NP - 1] = t0.y;
      }
      const float p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      if (PACKED) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint64_t pp = pack2(p[i], p[i]);
#pragma unroll
          for (int j = 0; j < NP; ++j) fma2(acc[i][j], pp, w[j]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            float2 a = *reinterpret_cast<float2*>(&acc[i][j]);
            const float2 ww = *reinterpret_cast<const float2*>(&w[j]);
            a.x = fmaf(p[i], ww.x, a.x);
            a.y = fmaf(p[i], ww.y, a.y);
            acc[i][j] = *reinterpret_cast<uint64_t*>(&a);
          }
        }
      }
    }
    __syncthreads();                              // buffer `buf` may be refilled (issued at the top of st+1)
  }
  // fold the odd-row accumulators into the even-row ones through shared memory (stage buffers are idle now)
  uint64_t* red = reinterpret_cast<uint64_t*>(psm);
  if (rg == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < NP; ++j) red[(i * NP + j) * 128 + (tid & 127)] = acc[i][j];
  }
  __syncthreads();
  if (rg == 1) return;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      float2 a = *reinterpret_cast<float2*>(&acc[i][j]);
      const float2 b = *reinterpret_cast<const float2*>(&red[(i * NP + j) * 128 + tid]);
      a.x += b.x; a.y += b.y;
      acc[i][j] = *reinterpret_cast<uint64_t*>(&a);
    }
  float* slab = out_partial + static_cast<int64_t>(blockIdx.y) * M * TP;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + cg * 4 + (i & 3) + (i >> 2) * 128;
    if (c < M) {
      uint64_t* dst = reinterpret_cast<uint64_t*>(slab + static_cast<int64_t>(c) * TP + tg * (2 * NP));
#pragma unroll
      for (int j = 0; j < NP; j += 2) *reinterpret_cast<ulonglong2*>(dst + j) = make_ulonglong2(acc[i][j], acc[i][j + 1]);
    }
  }
}

}  // namespace

int panel_splits(int64_t n_rows, int64_t M) {
  const int64_t tiles = (M + PC - 1) / PC;
  int64_t s = (148 * 2 + tiles - 1) / tiles;              // ~one full wave of 2 CTAs per SM
  const int64_t max_s = (n_rows + 4 * PR - 1) / (4 * PR); // at least 4 stages per CTA
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  // normalise: no empty split
  int64_t rps = ((n_rows + s - 1) / s + PR - 1) / PR * PR;
  s = (n_rows + rps - 1) / rps;
  return static_cast<int>(s);
}

int launch_panel_tmm(const float* P, int64_t ldp, const float* W, int64_t n_rows, int64_t M, int T_pad,
                     int n_splits, float* out_partial, cudaStream_t st) {
  if (n_rows <= 0 || M <= 0 || ldp < round_up(M, 128) || ldp % 4 != 0 || (reinterpret_cast<uintptr_t>(P) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(W) & 15) != 0)
    return set_error(ODF_ERR_ARG, "panel_tmm: bad shape (16-byte aligned P, W; ldp >= round_up(M,128), ldp % 4 == 0)");
  if (n_splits != panel_splits(n_rows, M)) return set_error(ODF_ERR_ARG, "panel_tmm: n_splits must come from odf_panel_splits");
  static DeviceOnce attr_once;
  bool& attr_set = attr_once.here();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(panel_tmm_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(panel_tmm_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(panel_tmm_kernel<32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(panel_tmm_kernel)");
    attr_set = true;
  }
  const int64_t rps = ((n_rows + n_splits - 1) / n_splits + PR - 1) / PR * PR;
  dim3 grid(static_cast<unsigned>((M + PC - 1) / PC), static_cast<unsigned>(n_splits));
  CUtensorMap tmP;
  int rc = make_map_plain_f32(&tmP, P, n_rows, ldp, ldp, PR, PC);
  if (rc) return rc;
  if (T_pad == 32 && getenv("ODF_PANEL_SCALAR"))
    panel_tmm_kernel<32, 0><<<grid, 256, PANEL_SMEM, st>>>(tmP, W, n_rows, static_cast<int>(M), rps, out_partial);
  else if (T_pad == 32)
    panel_tmm_kernel<32, 1><<<grid, 256, PANEL_SMEM, st>>>(tmP, W, n_rows, static_cast<int>(M), rps, out_partial);
  else if (T_pad == 16)
    panel_tmm_kernel<16, 1><<<grid, 256, PANEL_SMEM, st>>>(tmP, W, n_rows, static_cast<int>(M), rps, out_partial);
  else
    return set_error(ODF_ERR_ARG, "panel_tmm: T_pad must be 16 or 32");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "panel_tmm_kernel launch");
  return ODF_OK;
}

}  // namespace odf
