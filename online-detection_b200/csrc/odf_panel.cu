// Second half of a K^T (K v) sweep without recomputing K: the fused tile spills its K tiles
// (fp32, K_hi + K_lo exactly as the first contraction used them) into a transient row PANEL
// P [rows x ldp] and this kernel contracts it with the finished W = K v + w of the same rows:
//     out_partial[split][c][t] = sum_{r in split} P[r][c] * W[r][t]
// It replaces the second call of falkon `GaussianKernel.dmmv`'s inner product K_blk^T w
// (reached from InCoreFalkon.fit, FALKONWrapper_with_centers_selection_incore.py:68).
//
// The contraction has T <= 32 columns, i.e. 16 flop per panel byte: it is bound by streaming the
// panel from HBM once and by the fp32 FMA pipe, not by the tensor cores, so it is written as a
// register-tiled fp32 kernel (exact fp32 products, no operand split needed): 128 threads own a
// [128 centres x T_pad] accumulator tile (8 x T_pad/8 per thread), rows arrive through a cp.async
// double buffer, panel reads are 512-byte coalesced and conflict-free, W reads are broadcasts.
// Splits over rows write separate slabs that odf_finish_rows reduces in index order (deterministic).
#include "odf_internal.h"

namespace odf {
namespace {

constexpr int PC = 128;   // centres per CTA
constexpr int PR = 32;    // rows per stage
constexpr int PSTAGES = 2;   // 40 KB static smem -> 5 CTAs / SM keep ~200 KB of loads in flight per SM

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  const int bytes = valid ? 16 : 0;     // src-size 0 => the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int TP>
__global__ void __launch_bounds__(128)
panel_tmm_kernel(const float* __restrict__ P, int64_t ldp, const float* __restrict__ W, int64_t n_rows, int M,
                 int64_t rows_per_split, float* __restrict__ out_partial) {
  constexpr int TPT = TP / 8;                    // right-hand sides per thread (4 or 2)
  __shared__ __align__(16) float Ps[PSTAGES][PR][PC];
  __shared__ __align__(16) float Ws[PSTAGES][PR][TP];
  const int tid = threadIdx.x;
  const int cg = tid & 15, tg = tid >> 4;        // 16 centre groups of 8, 8 rhs groups of TPT
  const int c0 = blockIdx.x * PC;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.y) * rows_per_split;
  const int64_t r_end = min(n_rows, r_begin + rows_per_split);
  const int n_stages = static_cast<int>((r_end - r_begin + PR - 1) / PR);

  auto issue = [&](int st) {
    if (st < n_stages) {
      const int buf = st % PSTAGES;
      const int64_t r0 = r_begin + static_cast<int64_t>(st) * PR;
      // panel tile: PR rows x 512 B -> 32 x 32 chunks of 16 B, 8 per thread
#pragma unroll
      for (int i = 0; i < (PR * PC / 4) / 128; ++i) {
        const int idx = tid + i * 128;
        const int r = idx >> 5, ch = idx & 31;
        const bool ok = (r0 + r) < r_end;
        cp_async16(&Ps[buf][r][ch * 4], P + (ok ? (r0 + r) : r_begin) * ldp + c0 + ch * 4, ok);
      }
      // W tile: PR rows x TP floats
      for (int idx = tid; idx < PR * TP / 4; idx += 128) {
        const int r = idx / (TP / 4), ch = idx % (TP / 4);
        const bool ok = (r0 + r) < r_end;
        cp_async16(&Ws[buf][r][ch * 4], W + (ok ? (r0 + r) : r_begin) * TP + ch * 4, ok);
      }
    }
    cp_async_commit();
  };

  float acc[8][TPT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TPT; ++j) acc[i][j] = 0.f;

  for (int s = 0; s < PSTAGES - 1; ++s) issue(s);
  for (int st = 0; st < n_stages; ++st) {
    issue(st + PSTAGES - 1);
    cp_async_wait<PSTAGES - 1>();
    __syncthreads();
    const int buf = st % PSTAGES;
#pragma unroll 8
    for (int r = 0; r < PR; ++r) {
      const float4 p0 = *reinterpret_cast<const float4*>(&Ps[buf][r][cg * 8]);
      const float4 p1 = *reinterpret_cast<const float4*>(&Ps[buf][r][cg * 8 + 4]);
      float w[TPT];
      if (TPT == 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Ws[buf][r][tg * 4]);
        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[TPT - 1] = t.w;
      } else {
        const float2 t = *reinterpret_cast<const float2*>(&Ws[buf][r][tg * 2]);
        w[0] = t.x; w[TPT - 1] = t.y;
      }
      const float p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TPT; ++j) acc[i][j] = fmaf(p[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  float* slab = out_partial + static_cast<int64_t>(blockIdx.y) * M * TP;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + cg * 8 + i;
    if (c < M) {
      float* dst = slab + static_cast<int64_t>(c) * TP + tg * TPT;
      if (TPT == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][TPT - 1]);
      else *reinterpret_cast<float2*>(dst) = make_float2(acc[i][0], acc[i][TPT - 1]);
    }
  }
}

}  // namespace

int panel_splits(int64_t n_rows, int64_t M) {
  const int64_t tiles = (M + PC - 1) / PC;
  int64_t s = (148 * 10 + tiles - 1) / tiles;            // ~10 CTAs per SM worth of work items
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
  if (n_rows <= 0 || M <= 0 || ldp < round_up(M, PC) || ldp % 4 != 0)
    return set_error(ODF_ERR_ARG, "panel_tmm: bad shape (ldp must be >= round_up(M,128))");
  if (n_splits != panel_splits(n_rows, M)) return set_error(ODF_ERR_ARG, "panel_tmm: n_splits must come from odf_panel_splits");
  const int64_t rps = ((n_rows + n_splits - 1) / n_splits + PR - 1) / PR * PR;
  dim3 grid(static_cast<unsigned>((M + PC - 1) / PC), static_cast<unsigned>(n_splits));
  if (T_pad == 32)
    panel_tmm_kernel<32><<<grid, 128, 0, st>>>(P, ldp, W, n_rows, static_cast<int>(M), rps, out_partial);
  else if (T_pad == 16)
    panel_tmm_kernel<16><<<grid, 128, 0, st>>>(P, ldp, W, n_rows, static_cast<int>(M), rps, out_partial);
  else
    return set_error(ODF_ERR_ARG, "panel_tmm: T_pad must be 16 or 32");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "panel_tmm_kernel launch");
  return ODF_OK;
}

}  // namespace odf
