// HBM-bound helper kernels around the fused Gaussian tile: operand preparation (z-score +
// 3xTF32 split + norms), right-hand-side split/transposition, split-slab reduction and the
// per-column conjugate-gradient vector updates.  All are coalesced, vectorised where the layout
// allows it and reduction-order deterministic (no floating-point atomics).
#include <cuda_fp16.h>

#include "odf_ptx.cuh"
#include "odf_internal.h"

namespace odf {

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ prepare_points
// Pass A: one warp per row: |x'|^2 with x' = (x - mean) * scale, and the maximum over rows
// (non-negative floats order like their bit patterns, so an integer atomicMax does it).
template <bool VEC>
__global__ void __launch_bounds__(256)
rownorm_kernel(const float* __restrict__ X, int64_t n, int d, int64_t ldx,
               const float* __restrict__ mean, float scale, float* __restrict__ sqn, int64_t n_pad,
               unsigned int* __restrict__ maxbits) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  if (row >= n) {
    if (lane == 0) sqn[row] = 0.f;
    return;
  }
  const float* xr = X + row * ldx;
  double acc = 0.0;
  for (int c = lane * 4; c < d; c += 128) {
    float v[4];
    if (VEC && c + 3 < d) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(xr + c));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (c + k < d) ? __ldg(xr + c + k) : 0.f;
    }
    // four squares in fp32 (fma chain), the running sum in fp64: one conversion + one DADD per 4 elements instead of
    // 8 + 4 -- the fp64 pipe, not HBM, bound the first version (2.1 TB/s, ncu profiles/r2_ncu_vector_kernels.md); the
    // fp32 partial adds ~4e-7 absolute at |x|^2 = 400, far below the final rounding of the norm to fp32 (2.4e-5)
    float p4 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c + k < d) {
        float x = v[k];
        if (mean) x -= __ldg(mean + c + k);
        x *= scale;
        p4 = fmaf(x, x, p4);
      }
    }
    acc += static_cast<double>(p4);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const float f = static_cast<float>(acc);
    sqn[row] = f;
    atomicMax(maxbits, __float_as_uint(f));
  }
}

// Power-of-two operand scaling for fp16 storage: s * max|x| lands in (128, 256], so that the seed
// value s^2 |x|^2 / 2 stays below 2^15 and typical elements sit well inside fp16's normal range.
__device__ __forceinline__ float f16_operand_scale(unsigned int maxbits) {
  const float maxsq = __uint_as_float(maxbits);
  if (!(maxsq > 0.f) || !isfinite(maxsq)) return 1.f;
  float e = floorf(log2f(256.f * rsqrtf(maxsq)));
  e = fminf(fmaxf(e, -60.f), 60.f);
  return exp2f(e);
}

template <int KIND> struct OpElem;
template <> struct OpElem<KIND_TF32> {
  using type = float;
  static __device__ __forceinline__ void split(float v, float& hi, float& lo) { hi = tf32_rn(v); lo = tf32_rn(v - hi); }
};
template <> struct OpElem<KIND_F16> {
  using type = __half;
  static __device__ __forceinline__ void split(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
  }
};

// Pass B: one warp per row writes the split operand row:
//   [ features (d, zero padded to d_pad) | seed block: g_hi, g_lo, 0.. | ones block: 1, 1, 0.. ]
// with g = -s^2 |x'|^2 / 2 (the rank-1 accumulator seed of the tile kernel).  Only `hi` carries the
// two extra blocks.
template <int KIND, bool VEC, bool ZSEED = false>
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ X, int64_t n, int d, int64_t ldx, const float* __restrict__ mean,
             float scale, typename OpElem<KIND>::type* __restrict__ hi,
             typename OpElem<KIND>::type* __restrict__ lo, int d_pad, int pitch,
             const float* __restrict__ sqn, float* __restrict__ opscale) {
  using E = typename OpElem<KIND>::type;
  constexpr int BK = (KIND == KIND_F16) ? 64 : 32;
  const float s = (KIND == KIND_F16) ? f16_operand_scale(reinterpret_cast<const unsigned int*>(opscale)[1]) : 1.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) opscale[0] = s;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* xr = X + row * ldx;
  E* hr = hi + row * pitch;
  E* lr = lo + row * pitch;
  const float g = ZSEED ? 0.f : -0.5f * __ldg(sqn + row) * s * s;      // ZSEED: operands of the linear (GEMM) tile
  for (int c = lane * 4; c < pitch; c += 128) {
    float v[4];
    if (c < d_pad) {
      if (VEC && c + 3 < d) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(xr + c));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (c + k < d) ? __ldg(xr + c + k) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c + k < d) {
          if (mean) v[k] -= __ldg(mean + c + k);
          v[k] *= scale * s;
        } else {
          v[k] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = 0.f;
      if (c == d_pad + BK) { v[0] = 1.f; v[1] = 1.f; }           // ones block
    }
    E h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) OpElem<KIND>::split(v[k], h[k], l[k]);
    if (c == d_pad) {                                            // seed block: g_hi, g_lo
      E gh, gl;
      OpElem<KIND>::split(g, gh, gl);
      h[0] = gh; h[1] = gl;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) hr[c + k] = h[k];
    if (c < d_pad) {
#pragma unroll
      for (int k = 0; k < 4; ++k) lr[c + k] = l[k];
    }
  }
}

__global__ void __launch_bounds__(256)
zscore_kernel(float* __restrict__ X, int64_t n, int d, int64_t ldx, const float* __restrict__ mean,
              float scale) {
  const int64_t total = n * static_cast<int64_t>(d);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / d;
    const int c = static_cast<int>(i - r * d);
    float x = X[r * ldx + c];
    if (mean) x -= __ldg(mean + c);
    X[r * ldx + c] = x * scale;
  }
}

// ------------------------------------------------------------------ split_rhs
// V [m x T] -> vt_hi/vt_lo [T_pad x ldvt]; one block per 128 rows of V, transposed through smem.
__global__ void __launch_bounds__(128)
split_rhs_kernel(const float* __restrict__ V, int64_t m, int T, int64_t ldv, float scale,
                 float* __restrict__ vt_hi, float* __restrict__ vt_lo, int64_t ldvt, int T_pad) {
  __shared__ float tile[32][129];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 128;
  // coalesced-ish load: consecutive threads walk along a row of V
  for (int idx = threadIdx.x; idx < 128 * T_pad; idx += 128) {
    const int r = idx / T_pad, t = idx - r * T_pad;
    float v = 0.f;
    if (t < T && r0 + r < m) v = __ldg(V + (r0 + r) * ldv + t) * scale;
    tile[t][r] = v;
  }
  __syncthreads();
  const int r = threadIdx.x;
  if (r0 + r < ldvt) {
    for (int t = 0; t < T_pad; ++t) {
      const float v = tile[t][r];
      const float h = tf32_rn(v);
      vt_hi[t * ldvt + r0 + r] = h;
      vt_lo[t * ldvt + r0 + r] = tf32_rn(v - h);
    }
  }
}

// V [m x T] -> per-column max |scale V| (bits) for the fp16 split below
__global__ void __launch_bounds__(256)
col_absmax_kernel(const float* __restrict__ V, int64_t m, int T, int64_t ldv, float scale, uint32_t* __restrict__ absmax) {
  __shared__ float red[8][32];
  const int t = threadIdx.x & 31, w = threadIdx.x >> 5;
  float amax = 0.f;
  if (t < T)
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + w; r < m; r += static_cast<int64_t>(gridDim.x) * 8)
      amax = fmaxf(amax, fabsf(__ldg(V + r * ldv + t) * scale));
  red[w][t] = amax;
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) amax = fmaxf(amax, red[i][t]);
    if (t < T && amax > 0.f) atomicMax(absmax + t, __float_as_uint(amax));
  }
}

// V [m x T] -> vt_hi16 / vt_lo16 [T_pad x ldvt] fp16 (transposed): hi = rn16(s_t v), lo = rn16((s_t v - hi) 2^11) with the
// per-column power-of-two scale s_t = w16_scale(absmax[t]); the B operand of the pair tile's K.V contraction.
__global__ void __launch_bounds__(128)
split_rhs16_kernel(const float* __restrict__ V, int64_t m, int T, int64_t ldv, float scale, const uint32_t* __restrict__ absmax,
                   __half* __restrict__ vt_hi, __half* __restrict__ vt_lo, int64_t ldvt, int T_pad) {
  __shared__ float tile[32][129];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 128;
  for (int idx = threadIdx.x; idx < 128 * T_pad; idx += 128) {
    const int r = idx / T_pad, t = idx - r * T_pad;
    float v = 0.f;
    if (t < T && r0 + r < m) v = __ldg(V + (r0 + r) * ldv + t) * scale * w16_scale_from_bits(__ldg(absmax + t), false);
    tile[t][r] = v;
  }
  __syncthreads();
  const int r = threadIdx.x;
  if (r0 + r < ldvt) {
    for (int t = 0; t < T_pad; ++t) {
      const float v = tile[t][r];
      const __half h = __float2half_rn(v);
      vt_hi[t * ldvt + r0 + r] = h;
      vt_lo[t * ldvt + r0 + r] = __float2half_rn((v - __half2float(h)) * 2048.f);
    }
  }
}

// ------------------------------------------------------------------ finish (split-slab reduce)
// partial [S][n][T_pad] -> out[n x T]: each thread owns one row (a full 64/128-byte line per
// slab), sums the slabs in index order.
__global__ void __launch_bounds__(128)
finish_rows_kernel(const float* __restrict__ partial, int S, int64_t n, int T_pad, int T,
                   float scale, const float* __restrict__ addend, int64_t ld_add,
                   float* __restrict__ out, int64_t ldo, uint32_t* __restrict__ absmax) {
  __shared__ float tile[128][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 128;
  const int64_t r = r0 + threadIdx.x;
  float acc[32];
#pragma unroll
  for (int t = 0; t < 32; ++t) acc[t] = 0.f;
  if (r < n) {
    for (int s = 0; s < S; ++s) {
      const float4* src = reinterpret_cast<const float4*>(partial + (static_cast<int64_t>(s) * n + r) * T_pad);
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        if (4 * v < T_pad) {
          const float4 t4 = __ldg(src + v);
          acc[4 * v + 0] += t4.x; acc[4 * v + 1] += t4.y; acc[4 * v + 2] += t4.z; acc[4 * v + 3] += t4.w;
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 32; ++t) tile[threadIdx.x][t] = acc[t] * scale;
  __syncthreads();
  // coalesced store along rows of out
  for (int idx = threadIdx.x; idx < 128 * T; idx += 128) {
    const int rr = idx / T, t = idx - rr * T;
    float v = 0.f;
    if (r0 + rr < n) {
      v = tile[rr][t];
      if (addend) v += __ldg(addend + (r0 + rr) * ld_add + t);
      out[(r0 + rr) * ldo + t] = v;
    }
    if (absmax != nullptr) tile[rr][t] = fabsf(v);
  }
  if (absmax != nullptr) {                     // per-column max |out| (bit pattern of a non-negative float orders as uint)
    __syncthreads();
    if (threadIdx.x < T) {
      float amax = 0.f;
      for (int rr = 0; rr < 128; ++rr) amax = fmaxf(amax, tile[rr][threadIdx.x]);
      if (amax > 0.f) atomicMax(absmax + threadIdx.x, __float_as_uint(amax));
    }
  }
}

// W [n x ldw] fp32 -> W16 [round_up(n,128) x 64] fp16: columns 0..31 hi = rn16(s_t W), 32..63 lo = rn16((s_t W - hi) 2^11),
// s_t = w16_scale(max|W[:, t]|) a power of two per column; rows >= n and columns >= T are zero (the B operand of
// odf_panel16.cu).
__global__ void __launch_bounds__(256)
split_w16_kernel(const float* __restrict__ W, int64_t n, int64_t n_pad, int ldw, int T, const uint32_t* __restrict__ absmax,
                 __half* __restrict__ W16) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;     // one thread: one row, 8 columns
  const int64_t r = idx >> 2;
  const int t0 = static_cast<int>(idx & 3) * 8;
  if (r >= n_pad) return;
  __align__(16) __half hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int t = t0 + i;
    const float v = (r < n && t < T) ? __ldg(W + r * ldw + t) * w16_scale_from_bits(__ldg(absmax + t), false) : 0.f;
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn((v - __half2float(h)) * 2048.f);
  }
  *reinterpret_cast<uint4*>(W16 + r * 64 + t0) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(W16 + r * 64 + 32 + t0) = *reinterpret_cast<const uint4*>(lo);
}

__global__ void __launch_bounds__(128)
finish_split_kernel(const float* __restrict__ partial, int S, int64_t n, int T_pad, int T,
                    float scale, const float* __restrict__ addend, int64_t ld_add,
                    float* __restrict__ wt_hi, float* __restrict__ wt_lo, int64_t ldwt) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 128 + threadIdx.x;
  if (r >= ldwt) return;
  float acc[32];
#pragma unroll
  for (int t = 0; t < 32; ++t) acc[t] = 0.f;
  if (r < n) {
    for (int s = 0; s < S; ++s) {
      const float4* src = reinterpret_cast<const float4*>(partial + (static_cast<int64_t>(s) * n + r) * T_pad);
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        if (4 * v < T_pad) {
          const float4 t4 = __ldg(src + v);
          acc[4 * v + 0] += t4.x; acc[4 * v + 1] += t4.y; acc[4 * v + 2] += t4.z; acc[4 * v + 3] += t4.w;
        }
      }
    }
  }
  // consecutive threads -> consecutive columns of the transposed output: coalesced
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    if (t < T_pad) {
      float v = 0.f;
      if (r < n && t < T) {
        v = acc[t] * scale;
        if (addend) v += __ldg(addend + r * ld_add + t);
      }
      const float h = tf32_rn(v);
      wt_hi[t * ldwt + r] = h;
      wt_lo[t * ldwt + r] = tf32_rn(v - h);
    }
  }
}

// ------------------------------------------------------------------ CG column reductions
// Block (32 x 8): x walks the columns of a row (coalesced), y/blocks walk the rows.
// Stage 1 writes one double per (block, column); stage 2 sums the blocks in index order and
// applies the scalar recurrence.  Deterministic.
constexpr int RED_BLOCKS = 64;

template <int OP>  // 0: sum A*A   1: sum A*B
__global__ void __launch_bounds__(256)
colreduce_kernel(const float* __restrict__ A, const float* __restrict__ B, int64_t M, int T,
                 int64_t ld, double* __restrict__ part) {
  __shared__ double sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + tx;
    double acc = 0.0;
    if (t < T) {
      for (int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + ty; r < M; r += static_cast<int64_t>(gridDim.x) * 8) {
        const float a = __ldg(A + r * ld + t);
        const float b = OP == 0 ? a : __ldg(B + r * ld + t);
        acc += static_cast<double>(a) * static_cast<double>(b);
      }
    }
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && t < T) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += sm[k][tx];
      part[static_cast<int64_t>(blockIdx.x) * T + t] = s;
    }
    __syncthreads();
  }
}

// state layout: rs_old[T] | a[T] | b[T] | rs_new[T] | flag, pad[3]
__global__ void cg_init_final(const double* __restrict__ part, int nb, int T, float* state) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) {
    double s = 0.0;
    for (int k = 0; k < nb; ++k) s += part[static_cast<int64_t>(k) * T + t];
    state[t] = static_cast<float>(s);
    state[T + t] = 0.f;
    state[2 * T + t] = 0.f;
    state[3 * T + t] = static_cast<float>(s);
  }
  if (t == 0) {
    state[4 * T] = 0.f; state[4 * T + 1] = 0.f; state[4 * T + 2] = 0.f; state[4 * T + 3] = 0.f;
  }
}

__global__ void cg_alpha_final(const double* __restrict__ part, int nb, int T, float eps, float* state) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) {
    double s = 0.0;
    for (int k = 0; k < nb; ++k) s += part[static_cast<int64_t>(k) * T + t];
    const bool frozen = state[4 * T] != 0.f;
    const float pap = static_cast<float>(s);
    state[T + t] = frozen ? 0.f : state[t] / (pap + eps);
  }
}

// Single block: needs the max over columns before any column may update.
__global__ void cg_beta_final(const double* __restrict__ part, int nb, int T, float eps, float tol,
                              float* state) {
  __shared__ float smax[32];
  const bool frozen = state[4 * T] != 0.f;
  float mymax = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < nb; ++k) s += part[static_cast<int64_t>(k) * T + t];
    const float rs_new = static_cast<float>(s);
    if (!frozen) state[3 * T + t] = rs_new;
    mymax = fmaxf(mymax, fabsf(rs_new));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mymax = fmaxf(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mymax;
  __syncthreads();
  float m = 0.f;
  for (int k = 0; k < (blockDim.x >> 5); ++k) m = fmaxf(m, smax[k]);
  const bool conv = sqrtf(m) < tol;
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    if (frozen || conv) {
      state[2 * T + t] = 0.f;
    } else {
      const float rs_new = state[3 * T + t];
      state[2 * T + t] = rs_new / (state[t] + eps);
      state[t] = rs_new;
    }
  }
  if (threadIdx.x == 0 && conv && !frozen) state[4 * T] = 1.f;
}

// ------------------------------------------------------------------ CG elementwise updates
// mode 0: Y += sign * s[t] * X      (s = state + T: a)
// mode 1: Y  = X + s[t] * Y         (s = state + 2T: b)         [P = R + b P]
// mode 2: Y  = X - Z (unless frozen)                              [R = B - H]
template <int MODE>
__global__ void __launch_bounds__(256)
cg_elem_kernel(float* __restrict__ Y, const float* __restrict__ X, const float* __restrict__ Z,
               int64_t M, int T, int64_t ld, float sign, const float* __restrict__ state) {
  const bool frozen = state[4 * T] != 0.f;
  if (frozen) return;
  const int64_t total = M * static_cast<int64_t>(T);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / T;
    const int t = static_cast<int>(i - r * T);
    const int64_t o = r * ld + t;
    if (MODE == 0) {
      Y[o] = fmaf(sign * __ldg(state + T + t), __ldg(X + o), Y[o]);
    } else if (MODE == 1) {
      Y[o] = fmaf(__ldg(state + 2 * T + t), Y[o], __ldg(X + o));
    } else {
      Y[o] = __ldg(X + o) - __ldg(Z + o);
    }
  }
}

__global__ void __launch_bounds__(256)
axpby_kernel(float* __restrict__ out, float alpha, const float* __restrict__ A, float beta,
             const float* __restrict__ Bm, int64_t M, int T, int64_t ld) {
  const int64_t total = M * static_cast<int64_t>(T);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / T;
    const int64_t o = r * ld + (i - r * T);
    float v = alpha * A[o];
    if (Bm) v = fmaf(beta, Bm[o], v);
    out[o] = v;
  }
}

// ------------------------------------------------------------------ preconditioner helpers
__global__ void add_diag_kernel(float* __restrict__ A, int64_t M, float v) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < M) A[i * M + i] += v;
}
// Zero the strict lower triangle in row-major terms (== strict upper in cuBLAS column-major terms).
__global__ void __launch_bounds__(256)
zero_lower_kernel(float* __restrict__ A, int64_t M) {
  const int64_t r = blockIdx.y;
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < r;
       c += static_cast<int64_t>(gridDim.x) * blockDim.x)
    A[r * M + c] = 0.f;
}

__global__ void __launch_bounds__(256)
set_identity_kernel(float* __restrict__ A, int64_t M) {
  const int64_t total = M * M;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / M;
    A[i] = (i - r * M == r) ? 1.f : 0.f;
  }
}

int grid_for(int64_t total, int block, int cap = 148 * 8) {
  int64_t g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, what);
  return ODF_OK;
}

}  // namespace

// ---- internal C++ entry points used by odf_api.cu ------------------------------------------
int prepare_points(const float* X, int64_t n, int64_t d, int64_t ldx, const float* mean, float scale,
                   int kind, void* hi, void* lo, float* sqn, float* opscale, cudaStream_t st, bool zero_seed) {
  if (n <= 0 || d <= 0 || ldx < d) return set_error(ODF_ERR_ARG, "prepare_points: bad shape");
  if (kind != KIND_TF32 && kind != KIND_F16) return set_error(ODF_ERR_ARG, "prepare_points: unknown operand kind");
  const int64_t n_pad = round_up(n, 128);
  const int bk = static_cast<int>(kblock_elems(kind));
  const int d_pad = static_cast<int>(round_up(d, bk));
  const int pitch = d_pad + 2 * bk;
  const bool vec = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  cudaError_t e = cudaMemsetAsync(opscale, 0, 2 * sizeof(float), st);
  if (e != cudaSuccess) return set_cuda_error(e, "prepare_points: memset");
  unsigned int* maxbits = reinterpret_cast<unsigned int*>(opscale) + 1;
  const unsigned gridA = static_cast<unsigned>((n_pad + 7) / 8), gridB = static_cast<unsigned>((n + 7) / 8);
  const int di = static_cast<int>(d);
  if (vec) rownorm_kernel<true><<<gridA, 256, 0, st>>>(X, n, di, ldx, mean, scale, sqn, n_pad, maxbits);
  else rownorm_kernel<false><<<gridA, 256, 0, st>>>(X, n, di, ldx, mean, scale, sqn, n_pad, maxbits);
  if (zero_seed) {
    if (kind == KIND_F16) {
      if (vec) split_kernel<KIND_F16, true, true><<<gridB, 256, 0, st>>>(X, n, di, ldx, mean, scale, static_cast<__half*>(hi), static_cast<__half*>(lo), d_pad, pitch, sqn, opscale);
      else split_kernel<KIND_F16, false, true><<<gridB, 256, 0, st>>>(X, n, di, ldx, mean, scale, static_cast<__half*>(hi), static_cast<__half*>(lo), d_pad, pitch, sqn, opscale);
    } else {
      if (vec) split_kernel<KIND_TF32, true, true><<<gridB, 256, 0, st>>>(X, n, di, ldx, mean, scale, static_cast<float*>(hi), static_cast<float*>(lo), d_pad, pitch, sqn, opscale);
      else split_kernel<KIND_TF32, false, true><<<gridB, 256, 0, st>>>(X, n, di, ldx, mean, scale, static_cast<float*>(hi), static_cast<float*>(lo), d_pad, pitch, sqn, opscale);
    }
  } else if (kind == KIND_F16) {
    if (vec) split_kernel<KIND_F16, true><<<gridB, 256, 0, st>>>(X, n, di, ldx, mean, scale, static_cast<__half*>(hi), static_cast<__half*>(lo), d_pad, pitch, sqn, opscale);
    else split_kernel<KIND_F16, false><<<gridB, 256, 0, st>>>(X, n, di, ldx, mean, scale, static_cast<__half*>(hi), static_cast<__half*>(lo), d_pad, pitch, sqn, opscale);
  } else {
    if (vec) split_kernel<KIND_TF32, true><<<gridB, 256, 0, st>>>(X, n, di, ldx, mean, scale, static_cast<float*>(hi), static_cast<float*>(lo), d_pad, pitch, sqn, opscale);
    else split_kernel<KIND_TF32, false><<<gridB, 256, 0, st>>>(X, n, di, ldx, mean, scale, static_cast<float*>(hi), static_cast<float*>(lo), d_pad, pitch, sqn, opscale);
  }
  return check_launch("prepare kernels");
}

int zscore(float* X, int64_t n, int64_t d, int64_t ldx, const float* mean, float scale, cudaStream_t st) {
  if (n <= 0 || d <= 0) return set_error(ODF_ERR_ARG, "zscore: bad shape");
  zscore_kernel<<<grid_for(n * d, 256), 256, 0, st>>>(X, n, static_cast<int>(d), ldx, mean, scale);
  return check_launch("zscore_kernel");
}

int split_rhs(const float* V, int64_t m, int64_t T, int64_t ldv, float scale, float* vt_hi,
              float* vt_lo, int64_t ldvt, int T_pad, cudaStream_t st) {
  if (m <= 0 || T <= 0 || T > T_pad || T_pad > 32 || ldvt < m || ldvt % 128 != 0)
    return set_error(ODF_ERR_ARG, "split_rhs: bad shape");
  split_rhs_kernel<<<static_cast<unsigned>(ldvt / 128), 128, 0, st>>>(V, m, static_cast<int>(T), ldv, scale, vt_hi, vt_lo, ldvt, T_pad);
  return check_launch("split_rhs_kernel");
}

int split_rhs16(const float* V, int64_t m, int64_t T, int64_t ldv, float scale, uint32_t* absmax, void* vt_hi, void* vt_lo,
                int64_t ldvt, int T_pad, cudaStream_t st) {
  if (m <= 0 || T <= 0 || T > T_pad || T_pad > 32 || ldvt < m || ldvt % 128 != 0 || !absmax)
    return set_error(ODF_ERR_ARG, "split_rhs16: bad shape");
  cudaError_t e = cudaMemsetAsync(absmax, 0, 32 * sizeof(uint32_t), st);
  if (e != cudaSuccess) return set_cuda_error(e, "split_rhs16: memset");
  int64_t nb = (m + 63) / 64;
  if (nb > 592) nb = 592;
  col_absmax_kernel<<<static_cast<unsigned>(nb), 256, 0, st>>>(V, m, static_cast<int>(T), ldv, scale, absmax);
  split_rhs16_kernel<<<static_cast<unsigned>(ldvt / 128), 128, 0, st>>>(V, m, static_cast<int>(T), ldv, scale, absmax,
                                                                        static_cast<__half*>(vt_hi), static_cast<__half*>(vt_lo), ldvt, T_pad);
  return check_launch("split_rhs16");
}

int finish_rows(const float* partial, int S, int64_t n, int T_pad, int64_t T, float scale,
                const float* addend, int64_t ld_add, float* out, int64_t ldo, cudaStream_t st) {
  if (S <= 0 || n <= 0 || T <= 0 || T > T_pad || T_pad > 32) return set_error(ODF_ERR_ARG, "finish_rows: bad shape");
  finish_rows_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(partial, S, n, T_pad, static_cast<int>(T), scale, addend, ld_add, out, ldo, nullptr);
  return check_launch("finish_rows_kernel");
}

// W = sum of the partial slabs (+ addend) as finish_rows, then its fp16 hi/lo split for odf_panel16_tmm.
// Wf: fp32 scratch [n x T_pad]; absmax: 32 device words (per-column max |W| bits, reset here); W16: [round_up(n,128) x 64] fp16.
int finish_w16(const float* partial, int S, int64_t n, int T_pad, int64_t T, const float* addend, int64_t ld_add,
               float* Wf, uint32_t* absmax, void* W16, cudaStream_t st) {
  if (S <= 0 || n <= 0 || T <= 0 || T > T_pad || T_pad > 32 || !Wf || !absmax || !W16 ||
      (reinterpret_cast<uintptr_t>(W16) & 127) != 0)
    return set_error(ODF_ERR_ARG, "finish_w16: bad shape or alignment");
  cudaError_t e = cudaMemsetAsync(absmax, 0, 32 * sizeof(uint32_t), st);
  if (e != cudaSuccess) return set_cuda_error(e, "finish_w16: memset");
  finish_rows_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(partial, S, n, T_pad, static_cast<int>(T), 1.f, addend, ld_add, Wf, T_pad, absmax);
  const int64_t n_pad = (n + 127) / 128 * 128;
  split_w16_kernel<<<static_cast<unsigned>((n_pad * 4 + 255) / 256), 256, 0, st>>>(Wf, n, n_pad, T_pad, static_cast<int>(T), absmax, static_cast<__half*>(W16));
  return check_launch("finish_w16");
}

int finish_split(const float* partial, int S, int64_t n, int T_pad, int64_t T, float scale,
                 const float* addend, int64_t ld_add, float* wt_hi, float* wt_lo, int64_t ldwt,
                 cudaStream_t st) {
  if (S <= 0 || n <= 0 || T <= 0 || T > T_pad || T_pad > 32 || ldwt < n || ldwt % 128 != 0)
    return set_error(ODF_ERR_ARG, "finish_split: bad shape");
  finish_split_kernel<<<static_cast<unsigned>(ldwt / 128), 128, 0, st>>>(partial, S, n, T_pad, static_cast<int>(T), scale, addend, ld_add, wt_hi, wt_lo, ldwt);
  return check_launch("finish_split_kernel");
}

size_t cg_workspace_bytes(int64_t /*M*/, int64_t T) { return sizeof(double) * RED_BLOCKS * static_cast<size_t>(T); }

static int red_blocks(int64_t M) {
  int64_t b = (M + 7) / 8;
  if (b > RED_BLOCKS) b = RED_BLOCKS;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

int cg_init(const float* R, int64_t M, int64_t T, int64_t ld, float* state, void* ws, size_t wsb, cudaStream_t st) {
  if (wsb < cg_workspace_bytes(M, T)) return set_error(ODF_ERR_WORKSPACE, "cg_init: workspace too small");
  const int nb = red_blocks(M);
  colreduce_kernel<0><<<nb, 256, 0, st>>>(R, nullptr, M, static_cast<int>(T), ld, static_cast<double*>(ws));
  cg_init_final<<<static_cast<unsigned>((T + 127) / 128), 128, 0, st>>>(static_cast<double*>(ws), nb, static_cast<int>(T), state);
  return check_launch("cg_init");
}

int cg_alpha(const float* P, const float* AP, int64_t M, int64_t T, int64_t ld, float eps, float* state,
             void* ws, size_t wsb, cudaStream_t st) {
  if (wsb < cg_workspace_bytes(M, T)) return set_error(ODF_ERR_WORKSPACE, "cg_alpha: workspace too small");
  const int nb = red_blocks(M);
  colreduce_kernel<1><<<nb, 256, 0, st>>>(P, AP, M, static_cast<int>(T), ld, static_cast<double*>(ws));
  cg_alpha_final<<<static_cast<unsigned>((T + 127) / 128), 128, 0, st>>>(static_cast<double*>(ws), nb, static_cast<int>(T), eps, state);
  return check_launch("cg_alpha");
}

int cg_beta(const float* R, int64_t M, int64_t T, int64_t ld, float eps, float tol, float* state,
            void* ws, size_t wsb, cudaStream_t st) {
  if (wsb < cg_workspace_bytes(M, T)) return set_error(ODF_ERR_WORKSPACE, "cg_beta: workspace too small");
  const int nb = red_blocks(M);
  colreduce_kernel<0><<<nb, 256, 0, st>>>(R, nullptr, M, static_cast<int>(T), ld, static_cast<double*>(ws));
  cg_beta_final<<<1, 128, 0, st>>>(static_cast<double*>(ws), nb, static_cast<int>(T), eps, tol, state);
  return check_launch("cg_beta");
}

int cg_axpy_a(float* Y, const float* X, int64_t M, int64_t T, int64_t ld, float sign, const float* state, cudaStream_t st) {
  cg_elem_kernel<0><<<grid_for(M * T, 256), 256, 0, st>>>(Y, X, nullptr, M, static_cast<int>(T), ld, sign, state);
  return check_launch("cg_axpy_a");
}
int cg_xpby_b(float* P, const float* R, int64_t M, int64_t T, int64_t ld, const float* state, cudaStream_t st) {
  cg_elem_kernel<1><<<grid_for(M * T, 256), 256, 0, st>>>(P, R, nullptr, M, static_cast<int>(T), ld, 1.f, state);
  return check_launch("cg_xpby_b");
}
int cg_residual(float* R, const float* Bm, const float* H, int64_t M, int64_t T, int64_t ld, const float* state, cudaStream_t st) {
  cg_elem_kernel<2><<<grid_for(M * T, 256), 256, 0, st>>>(R, Bm, H, M, static_cast<int>(T), ld, 1.f, state);
  return check_launch("cg_residual");
}
int axpby(float* out, float alpha, const float* A, float beta, const float* Bm, int64_t M, int64_t T, int64_t ld, cudaStream_t st) {
  axpby_kernel<<<grid_for(M * T, 256), 256, 0, st>>>(out, alpha, A, beta, Bm, M, static_cast<int>(T), ld);
  return check_launch("axpby");
}
int add_diag(float* A, int64_t M, float v, cudaStream_t st) {
  add_diag_kernel<<<static_cast<unsigned>((M + 255) / 256), 256, 0, st>>>(A, M, v);
  return check_launch("add_diag");
}
int set_identity(float* A, int64_t M, cudaStream_t st) {
  set_identity_kernel<<<grid_for(M * M, 256, 148 * 16), 256, 0, st>>>(A, M);
  return check_launch("set_identity");
}
int zero_lower(float* A, int64_t M, cudaStream_t st) {
  dim3 grid(8, static_cast<unsigned>(M));
  zero_lower_kernel<<<grid, 256, 0, st>>>(A, M);
  return check_launch("zero_lower");
}

}  // namespace odf
