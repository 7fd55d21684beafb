// Application of the explicit inverse factors of the FALKON preconditioner inside the CG loop (falkon
// `FalkonPreconditioner.invT / invTt / invA / invAt`; four per CG iteration):
//     out = U B      or      out = U^T B,        U upper triangular M x M (fp32),  B: M x T (T <= 32, fp32)
// Round 1 ran these as cuBLAS sgemm on the full square (0.175 ms at M = 10 k: 400 MB read, half of it zeros, SIMT fp32
// accumulation).  This kernel reads only the triangle and accumulates in FP64: the products are exact (fp32 x fp32 in
// fp64) and the sums carry no fp32 rounding, so an application is an exactly rounded linear map of its fp32 inputs --
// the applications were a main source of the arithmetic noise that 20 CG iterations amplify into the scores
// (profiles/r2_accuracy_probe_200k_4k.log: triangular solves 1.7e-4, explicit inverse through sgemm 3.7e-4 against the fp64
// oracle).  Bound: M^2 T / 2 fp64 FMAs (1.6e9 at M = 10 k, T = 32) against 200 MB of triangle (31 us of HBM): bound by the
// fp64 rate, which on this part saturates at 17 TFLOP/s (0.19 ms) whichever way it is issued -- the time of the library's
// fp32 call, with fp64 accumulation and half the bytes.
//
// Persistent CTAs (one per SM) deal the 32-row output blocks out snake-wise, largest first (block b of a round goes to CTA b
// in even rounds and to CTA G-1-b in odd ones): a block's work is proportional to its part of the triangle, so two
// consecutive rounds give every CTA the same number of columns and nothing has to be split over CTAs (no partial
// slabs, no atomics: deterministic).  History (M = 10 k, T = 30, profiles/r2_tri_apply.md): v1 launched ceil(blocks / 2)
// CTAs of one block pair each -- 157 CTAs on 148 SMs, a second wave of 9 CTAs doubled the time -- and fed scalar DFMAs
// with 17 shared-memory loads per 32 of them: 0.57 ms, fp64 pipe 30 % active.  v2 (persistent, 4 x 8 register blocks, 6
// 128-bit loads per 32 DFMAs) was still bound by the shared-memory port (a 128-bit load costs 4 wavefronts however many
// lanes share an address): 0.31 ms.  This version contracts on the FP64 tensor cores, mma.sync m8n8k4: one load per lane
// and operand fragment, 8 loads per 16 DMMAs = 4096 FMAs: 0.19 ms = 17 TFLOP/s, the DMMA pipe's limit (see dmma884).
// 512 threads: the 16 warps split every 64-wide slice of the contraction (one k = 4 step each); every warp holds the whole
// 32 x 32 output block as 4 x 4 accumulator fragments (32 fp64 registers per lane).  The fp32 U tile and B slice of a
// slice travel global -> shared memory as 4-byte cp.async copies (zero-filled outside the triangle / the matrix) through
// a ring of 4 stages: three slices (48 KB per SM) are in flight while one is contracted.  U is staged as [row][c] for the
// plain and [c][row] for the transposed orientation: the global reads are whole 128-byte lines in both, and pitches of
// 4 resp. 8 mod 32 words make the fragment loads conflict-free; the values are widened to fp64 in registers (exact).
// The partial sums of the warps are added in warp order at the end of a block.
#include "odf_internal.h"

namespace odf {

namespace {

constexpr int TR = 32;      // output rows per block
constexpr int TC = 64;      // contraction slice
constexpr int TW = 16;      // warps per CTA: TC / TW = 4 contraction values (one m8n8k4 step) per warp and slice
constexpr int TNS = 4;      // ring stages
constexpr int PU0 = 68;     // plain:      Us[r * PU0 + c]   (32 rows x 64 values)
constexpr int PU1 = 40;     // transposed: Us[c * PU1 + r]   (64 values x 32 rows)
constexpr int PB = 40;      //             Bs[c * PB + t]    (64 values x 32 right-hand sides)
constexpr int US = (TR * PU0 > TC * PU1) ? TR * PU0 : TC * PU1;
constexpr int TSTAGE = US + TC * PB;                       // floats per stage
constexpr int TRI_SMEM = TNS * TSTAGE * 4;                 // 81920 B; the reduction buffer (32 KB) aliases it
static_assert(TRI_SMEM >= 4 * 16 * 32 * 16, "reduction buffer");

// D (8 x 8) += A (8 x 4, row) . B (4 x 8, col): lane l holds A[l / 4][l % 4], B[l % 4][l / 4], D[l / 4][2 (l % 4) + {0, 1}].
// (DMMA.8x8x4 is the native shape of sm_100a: ptxas lowers m16n8k4 to pairs of it.  ncu shows the DMMA sub-pipe 50 % busy
// with math-pipe throttle as the top stall reason, at 17 TFLOP/s: the FP64 ceiling of this kernel on a B200.)
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// 4-byte asynchronous copy, zero-filled when !valid (the source is not read then)
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// NJ = column fragments of 8 right-hand sides a warp accumulates: 4 (T <= 32) or 1 (T <= 8: the one-class fits of the
// minibootstrap, where the kernel is bound by the latency of a block's slice chain, not by the DMMA rate)
template <int TRANSPOSED, int NJ>
__global__ void __launch_bounds__(32 * TW, 1)
tri_apply_kernel(const float* __restrict__ U, int64_t M, const float* __restrict__ Bm, int64_t ldb, int T,
                 float* __restrict__ out, int64_t ldo, int64_t row0, int64_t row1, int n_blocks) {
  extern __shared__ __align__(16) float tri_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lq = lane >> 2, lr = lane & 3;
  const int G = static_cast<int>(gridDim.x);
  for (int round = 0;; ++round) {
    const int idx = round * G + ((round & 1) ? (G - 1 - static_cast<int>(blockIdx.x)) : static_cast<int>(blockIdx.x));
    if (idx >= n_blocks) break;
    const int blk = TRANSPOSED ? (n_blocks - 1 - idx) : idx;               // largest part of the triangle first
    const int64_t r0 = row0 + static_cast<int64_t>(blk) * TR;
    const int rn = static_cast<int>((row1 - r0 < TR) ? (row1 - r0) : TR);
    // contraction range of these rows: plain c in [r0, M) (U[r][c] = 0 for c < r); transposed c in [0, r0 + rn)
    const int64_t cbeg = TRANSPOSED ? 0 : (r0 / TC) * TC;
    const int64_t cend = TRANSPOSED ? (r0 + rn) : M;
    const int n_slices = static_cast<int>((cend - cbeg + TC - 1) / TC);
    double acc[4][NJ][2];                    // [row tile i][column tile j]: rows 8 i + lq, columns 8 j + 2 lr, + 1
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // slice sl -> stage sl % TNS (one commit group per slice, empty past the end)
    auto issue = [&](int sl) {
      if (sl < n_slices) {
        const int64_t c0 = cbeg + static_cast<int64_t>(sl) * TC;
        float* Us = tri_smem + (sl % TNS) * TSTAGE;
        float* Bs = Us + US;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (!TRANSPOSED) {
            // warp -> rows 2 w, 2 w + 1, lane -> column (two 32-wide halves): coalesced 128-byte lines
            const int rl = 2 * warp + (k >> 1);
            const int64_t rr = r0 + rl, cgl = c0 + (k & 1) * 32 + lane;
            const bool ok = rl < rn && cgl < M && cgl >= rr;
            cp_async4(Us + rl * PU0 + (k & 1) * 32 + lane, ok ? U + rr * M + cgl : U, ok);
          } else {
            // warp -> rows 4 w .. 4 w + 3 of U (contraction index), lane -> column r: coalesced
            const int64_t cgl = c0 + 4 * warp + k, rgl = r0 + lane;
            const bool ok = cgl < cend && lane < rn && rgl >= cgl;
            cp_async4(Us + (4 * warp + k) * PU1 + lane, ok ? U + cgl * M + rgl : U, ok);
          }
          const int64_t cb = c0 + 4 * warp + k;
          const bool okb = cb < cend && lane < T;
          cp_async4(Bs + (4 * warp + k) * PB + lane, okb ? Bm + cb * ldb + lane : Bm, okb);
        }
      }
      cp_async_commit();
    };
    __syncthreads();                         // the previous block's reduction buffer has been consumed
#pragma unroll
    for (int sl = 0; sl < TNS - 1; ++sl) issue(sl);
    for (int sl = 0; sl < n_slices; ++sl) {
      cp_async_wait<TNS - 2>();              // this thread's copies of slice sl have landed ...
      __syncthreads();                       // ... everybody's have, and slice sl - 1 has been consumed by every warp
      issue(sl + TNS - 1);                   // into the stage of slice sl - 1
      const float* Us = tri_smem + (sl % TNS) * TSTAGE;
      const float* Bs = Us + US;
      const int c = 4 * warp + lr;           // this lane's contraction value of the warp's k = 4 step
      double a[4], b[NJ];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = static_cast<double>(TRANSPOSED ? Us[c * PU1 + 8 * i + lq] : Us[(8 * i + lq) * PU0 + c]);
#pragma unroll
      for (int j = 0; j < NJ; ++j) b[j] = static_cast<double>(Bs[c * PB + 8 * j + lq]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    cp_async_wait<0>();
    // cross-warp reduction in warp order, 4 warps at a time: red[w][s = 4 i + j][lane] (double2) aliases the ring
    double2* red = reinterpret_cast<double2*>(tri_smem);
    double2 total = make_double2(0.0, 0.0);
#pragma unroll 1
    for (int pass = 0; pass < TW / 4; ++pass) {
      __syncthreads();
      if ((warp >> 2) == pass) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) red[((warp & 3) * 16 + 4 * i + j) * 32 + lane] = make_double2(acc[i][j][0], acc[i][j][1]);
      }
      __syncthreads();
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const double2 v = red[w * 512 + threadIdx.x];
        total.x += v.x;
        total.y += v.y;
      }
    }
    {
      // thread -> (s = 4 i + j, lane') of the layout above: row 8 i + lq', columns 8 j + 2 lr', + 1
      const int s = threadIdx.x >> 5, i = s >> 2, j = s & 3;
      const int r = 8 * i + lq, t = 8 * j + 2 * lr;
      if (r < rn && j < NJ) {
        float* o = out + (r0 - row0 + r) * ldo + t;
        if (t < T) o[0] = static_cast<float>(total.x);
        if (t + 1 < T) o[1] = static_cast<float>(total.y);
      }
    }
  }
}

int tri_sm_count() { return device_sm_count(); }

}  // namespace

// out[(r - row0), :] = (op(U) B)[r, :] for r in [row0, row1); out has pitch ldo and row1 - row0 rows
int tri_apply(const float* U, int64_t M, const float* Bm, int64_t ldb, int64_t T, float* out, int64_t ldo, int64_t row0, int64_t row1,
              int transposed, cudaStream_t st) {
  if (M <= 0 || T <= 0 || T > 32 || row0 < 0 || row1 > M || row0 >= row1 || ldb < T || ldo < T)
    return set_error(ODF_ERR_ARG, "tri_apply: bad shape (T <= 32 per call)");
  const int n_blocks = static_cast<int>((row1 - row0 + TR - 1) / TR);
  static DeviceOnce attr_once;
  bool& attr_set = attr_once.here();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tri_apply_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tri_apply_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tri_apply_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tri_apply_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_SMEM);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(tri_apply_kernel)");
    attr_set = true;
  }
  const int sms = tri_sm_count();
  const int grid = n_blocks < sms ? n_blocks : sms;
  const int Ti = static_cast<int>(T);
  if (T <= 8) {
    if (transposed) tri_apply_kernel<1, 1><<<grid, 32 * TW, TRI_SMEM, st>>>(U, M, Bm, ldb, Ti, out, ldo, row0, row1, n_blocks);
    else tri_apply_kernel<0, 1><<<grid, 32 * TW, TRI_SMEM, st>>>(U, M, Bm, ldb, Ti, out, ldo, row0, row1, n_blocks);
  } else {
    if (transposed) tri_apply_kernel<1, 4><<<grid, 32 * TW, TRI_SMEM, st>>>(U, M, Bm, ldb, Ti, out, ldo, row0, row1, n_blocks);
    else tri_apply_kernel<0, 4><<<grid, 32 * TW, TRI_SMEM, st>>>(U, M, Bm, ldb, Ti, out, ldo, row0, row1, n_blocks);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ODF_OK : set_cuda_error(e, "tri_apply_kernel launch");
}

}  // namespace odf
