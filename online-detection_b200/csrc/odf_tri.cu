// Application of the explicit inverse factors of the FALKON preconditioner inside the CG loop (falkon
// `FalkonPreconditioner.invT / invTt / invA / invAt`; four per CG iteration):
//     out = U B      or      out = U^T B,        U upper triangular M x M (fp32),  B: M x T (T <= 32, fp32)
// Round 1 ran these as cuBLAS sgemm on the full square (0.175 ms at M = 10 k: 400 MB read, half of it zeros, SIMT fp32
// accumulation).  This kernel reads only the triangle and accumulates in FP64: the products are exact (fp32 x fp32 in
// fp64) and the sums carry no fp32 rounding, so an application is an exactly rounded linear map of its fp32 inputs --
// the applications were a main source of the arithmetic noise that 20 CG iterations amplify into the scores
// (profiles/r2_accuracy_probe_200k_4k.log: triangular solves 1.7e-4, explicit inverse through sgemm 3.7e-4 against the fp64
// oracle).  Bound: M^2 T / 2 fp64 FMAs (1.6e9 at M = 10 k, T = 32: ~90 us at the B200's 64 DFMA / clk / SM) against
// 200 MB of triangle (31 us of HBM): DFMA-bound, and still half the time of the library call.
//
// One CTA owns a PAIR of 32-row output blocks (i, nb-1-i): together they see M + 32 columns of the triangle whatever i is,
// so the CTAs are balanced without splitting a row's sum over CTAs (no partial slabs, no atomics: deterministic).
// 256 threads: lane l holds output row l of the block and all 32 right-hand sides in registers (32 fp64 accumulators); the
// 8 warps split every 64-wide slice of the contraction (8 values each) and their partial sums are added in warp order at
// the end of the block.  Per contraction value a warp does one conflict-free 64-bit shared-memory load of its U column and
// 16 broadcast 128-bit loads of the B row for 32 DFMAs per lane: the fp64 pipe, not the shared-memory port, sets the pace
// (a first version with the right-hand sides across the lanes needed one shared-memory wavefront per DFMA).  The U tile
// and the B slice are widened to fp64 once, while they are staged; global loads are whole 128-byte lines in both
// orientations and the loads of slice i + 1 are in flight during the FMAs of slice i.
#include "odf_internal.h"

namespace odf {

namespace {

constexpr int TR = 32;     // output rows per block (one per lane)
constexpr int TC = 64;     // contraction slice (8 values per warp)
constexpr int UP = 33;     // pitch of the U tile in doubles: the transposing stores of the plain orientation stay 2-way

template <int TRANSPOSED>
__global__ void __launch_bounds__(256)
tri_apply_kernel(const float* __restrict__ U, int64_t M, const float* __restrict__ Bm, int64_t ldb, int T,
                 float* __restrict__ out, int64_t ldo, int64_t row0, int64_t row1, int n_blocks) {
  // staging area [Us | Bs], reused for the cross-warp reduction at the end of a row block
  __shared__ __align__(16) double smem[TC * UP + TC * 32];
  double (*Us)[UP] = reinterpret_cast<double (*)[UP]>(smem);              // Us[c][r]
  double (*Bs)[32] = reinterpret_cast<double (*)[32]>(smem + TC * UP);    // Bs[c][t]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int half = 0; half < 2; ++half) {
    const int blk = half == 0 ? static_cast<int>(blockIdx.x) : n_blocks - 1 - static_cast<int>(blockIdx.x);
    if (half == 1 && blk <= static_cast<int>(blockIdx.x)) break;           // odd count: the middle block is done once
    const int64_t r0 = row0 + static_cast<int64_t>(blk) * TR;
    const int64_t rn = (row1 - r0 < TR) ? (row1 - r0) : TR;
    // contraction range of these rows: plain c in [r0, M) (U[r][c] = 0 for c < r); transposed c in [0, r0 + rn)
    const int64_t cbeg = TRANSPOSED ? 0 : (r0 / TC) * TC;
    const int64_t cend = TRANSPOSED ? (r0 + rn) : M;
    double acc[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) acc[t] = 0.0;
    float ru[8], rb[8];
    auto fetch = [&](int64_t c0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = 0.f;
        if (!TRANSPOSED) {
          // warp -> rows 4 w .. 4 w + 3, lane -> column (two 32-wide halves): coalesced 128-byte lines
          const int64_t rr = r0 + 4 * warp + (k >> 1), cg = c0 + (k & 1) * 32 + lane;
          if (4 * warp + (k >> 1) < rn && cg < M && cg >= rr) v = __ldg(U + rr * M + cg);
        } else {
          // warp -> rows 8 w .. 8 w + 7 of U (contraction index), lane -> column r: coalesced
          const int64_t cg = c0 + 8 * warp + k, rg = r0 + lane;
          if (cg < cend && lane < rn && rg >= cg) v = __ldg(U + cg * M + rg);
        }
        ru[k] = v;
        const int64_t cb = c0 + 8 * warp + k;
        rb[k] = (cb < cend && lane < T) ? __ldg(Bm + cb * ldb + lane) : 0.f;
      }
    };
    fetch(cbeg);
    for (int64_t c0 = cbeg; c0 < cend; c0 += TC) {
      __syncthreads();                       // the previous slice (or the previous block's reduction) has been consumed
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!TRANSPOSED) Us[(k & 1) * 32 + lane][4 * warp + (k >> 1)] = static_cast<double>(ru[k]);
        else Us[8 * warp + k][lane] = static_cast<double>(ru[k]);
        Bs[8 * warp + k][lane] = static_cast<double>(rb[k]);
      }
      __syncthreads();
      if (c0 + TC < cend) fetch(c0 + TC);
#pragma unroll 2
      for (int k = 0; k < 8; ++k) {
        const int c = 8 * warp + k;
        const double u = Us[c][lane];
        const double2* brow = reinterpret_cast<const double2*>(&Bs[c][0]);
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const double2 b = brow[t];
          acc[2 * t] = fma(u, b.x, acc[2 * t]);
          acc[2 * t + 1] = fma(u, b.y, acc[2 * t + 1]);
        }
      }
    }
    // cross-warp reduction in warp order, 16 right-hand sides at a time: red[w][r][16] aliases the staging area
    double (*red)[TR][16] = reinterpret_cast<double (*)[TR][16]>(smem);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      __syncthreads();
#pragma unroll
      for (int t = 0; t < 16; ++t) red[warp][lane][t] = acc[16 * h + t];
      __syncthreads();
      // 32 rows x 16 columns = 512 sums over 256 threads
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int idx = threadIdx.x + 256 * e;
        const int r = idx >> 4, t = idx & 15;
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[w][r][t];
        if (r < rn && 16 * h + t < T) out[(r0 - row0 + r) * ldo + 16 * h + t] = static_cast<float>(sum);
      }
    }
  }
}

}  // namespace

// out[(r - row0), :] = (op(U) B)[r, :] for r in [row0, row1); out has pitch ldo and row1 - row0 rows
int tri_apply(const float* U, int64_t M, const float* Bm, int64_t ldb, int64_t T, float* out, int64_t ldo, int64_t row0, int64_t row1,
              int transposed, cudaStream_t st) {
  if (M <= 0 || T <= 0 || T > 32 || row0 < 0 || row1 > M || row0 >= row1 || ldb < T || ldo < T)
    return set_error(ODF_ERR_ARG, "tri_apply: bad shape (T <= 32 per call)");
  const int n_blocks = static_cast<int>((row1 - row0 + TR - 1) / TR);
  const int grid = (n_blocks + 1) / 2;
  if (transposed) tri_apply_kernel<1><<<grid, 256, 0, st>>>(U, M, Bm, ldb, static_cast<int>(T), out, ldo, row0, row1, n_blocks);
  else tri_apply_kernel<0><<<grid, 256, 0, st>>>(U, M, Bm, ldb, static_cast<int>(T), out, ldo, row0, row1, n_blocks);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ODF_OK : set_cuda_error(e, "tri_apply_kernel launch");
}

}  // namespace odf
