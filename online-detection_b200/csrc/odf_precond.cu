// Tensor-core build of the FALKON preconditioner (falkon `FalkonPreconditioner.init`; reached from
// InCoreFalkon.fit, src/modules/region-classifier/FALKONWrapper_with_centers_selection_incore.py:68):
//
//     T = chol_upper(K_MM + eps M I)        A = chol_upper(T T^T / M + lam I)        (+ explicit T^-1, A^-1)
//
// odf_precond_init hands the two factorisations to cuSOLVER potrf (SIMT fp32, ~32 TFLOP/s at M = 10 k) and every product
// to cuBLAS sgemm (~60 TFLOP/s): 67 ms at M = 10 k, 18 % of a one-GPU fit and more than half of an eight-GPU one.  Here
// every O(M^3) flop is a 3-pass split-fp16 product on tcgen05 (the LINEAR store variant of the fused tile,
// odf_gauss_tile.cu: ~240 TFLOP/s algorithmic) and the library only factorises / solves against the nb x nb diagonal
// blocks.  Everything is written for row-major LOWER factors, because then every product is an "NT" GEMM
// C = A B^T with both operands K-major -- the only form the tile consumes:
//
//   blocked right-looking Cholesky  G = L L^T      diagonal block     cusolverDnSpotrf (nb = 1024)
//                                                  row panel          L21 = G21 L11^-T          cublasStrsm
//                                                  trailing update    G22 -= L21 L21^T          tile, k = nb, block columns
//   T = L^T                                        in-place transpose (lower -> upper, strict lower zeroed)
//   T T^T / M + lam I (lower blocks only)          k-slices of T: C[J0:k1, J] += T[J0:k1, K] T[J, K]^T / M      tile
//   L^-1 (divide and conquer, leaves nb)           leaf               cublasStrsm against the identity (side streams)
//                                                  X21 = -X22 (X11^T L21^T)^T: two NT GEMMs per node, k-slices see only
//                                                  the non-zero rows of the triangular operand             tile
//
// The tensor core adds with truncation, so no contraction chain is longer than KSLICE = 1024 (slices are summed by the
// epilogue in fp32 round-to-nearest, beta = 1); tools/precision_study.py B sized that.  Accuracy: ~1e-6 of |A||B|^T per
// product (sgemm: 3e-7).  A perturbed (T, A) pair that is used consistently leaves the fixed point of the preconditioned
// system unchanged up to the regulariser lam T^T T, so this only has to be a good preconditioner -- but a failed pivot
// (ODF_ERR_LINALG) is reported exactly like the library path does, and the caller can fall back to it.
#include <cstdlib>
#include <vector>

#include <cublas_v2.h>
#include <cusolverDn.h>

#include "odf_internal.h"

namespace odf {

// implemented in odf_api.cu / odf_vec.cu
int lib_handles(cudaStream_t st, cublasHandle_t* blas, cusolverDnHandle_t* solver);
int prepare_points(const float*, int64_t, int64_t, int64_t, const float*, float, int, void*, void*, float*, float*, cudaStream_t,
                   bool zero_seed);
int add_diag(float*, int64_t, float, cudaStream_t);

namespace {

constexpr int64_t NB = 1024;       // block column of the tensor-core updates, leaf of the triangular inverse
constexpr int64_t KSLICE = 1024;   // longest tensor-core accumulation chain
constexpr int N_SIDE = 4;          // side streams for the independent leaf inversions
constexpr int64_t NBC_MAX = 4096;  // largest diagonal block handed to cuSOLVER

// Diagonal block of the blocked Cholesky (ODF_PRECOND_NB = 1024 / 2048 / 4096).  cuSOLVER's own potrf is latency-bound at
// these sizes but much better per column than a chain of small library calls (measured on B200: potrf 0.35 / 0.67 /
// 1.43 ms at n = 1024 / 2048 / 4096, trsm against a 1024-wide triangle 0.5 ms): fewer, larger diagonal blocks shorten
// the sequential chain, everything off the diagonal still runs on the tensor cores.
int64_t chol_block() {
  static int64_t nbc = 0;
  if (nbc == 0) {
    const char* e = getenv("ODF_PRECOND_NB");
    int64_t v = e ? atoll(e) : 2048;
    if (v < NB) v = NB;
    if (v > NBC_MAX) v = NBC_MAX;
    nbc = v / NB * NB;
  }
  return nbc;
}
bool potrf_via_copy() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ODF_PRECOND_POTRF_COPY");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

inline size_t al256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }

// ---- small kernels ---------------------------------------------------------------------------------------------------
// In place, n x n with pitch ld: upper <- (lower)^T, strict lower <- 0.  One CTA per 32 x 32 tile pair (bi >= bj).
__global__ void __launch_bounds__(256)
tril_to_triu_kernel(float* __restrict__ A, int64_t n, int64_t ld, int nt) {
  __shared__ float tile[32][33];
  // linear index -> (bi, bj) with bi >= bj
  int bi = static_cast<int>((sqrtf(8.f * static_cast<float>(blockIdx.x) + 1.f) - 1.f) * 0.5f);
  while (static_cast<int64_t>(bi + 1) * (bi + 2) / 2 <= blockIdx.x) ++bi;
  while (static_cast<int64_t>(bi) * (bi + 1) / 2 > blockIdx.x) --bi;
  const int bj = static_cast<int>(blockIdx.x - static_cast<int64_t>(bi) * (bi + 1) / 2);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  const int64_t r0 = static_cast<int64_t>(bi) * 32, c0 = static_cast<int64_t>(bj) * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = r0 + ty + 8 * k, c = c0 + tx;
    tile[ty + 8 * k][tx] = (r < n && c < n && c <= r) ? A[r * ld + c] : 0.f;
  }
  __syncthreads();
  // lower tile (bi, bj): zero (strict lower only on the diagonal tile)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = r0 + ty + 8 * k, c = c0 + tx;
    if (r < n && c < n && c < r) A[r * ld + c] = 0.f;
  }
  // upper tile (bj, bi): transposed
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = c0 + ty + 8 * k, c = r0 + tx;             // element (r, c) of the upper tile = lower (c, r)
    if (r < n && c < n && c >= r && (bi != bj || c > r)) A[r * ld + c] = tile[tx][ty + 8 * k];
  }
  (void)nt;
}

// out (upper, strict lower zeroed) = (lower triangle of in)^T, both n x n with pitch ld; one CTA per 32 x 32 tile of out
__global__ void __launch_bounds__(256)
tril_transpose_copy_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, int64_t ld) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int by = blockIdx.y, bx = blockIdx.x;                     // tile (by, bx) of out
  const int64_t r0 = static_cast<int64_t>(by) * 32, c0 = static_cast<int64_t>(bx) * 32;
  if (bx >= by) {
    // tile (bx, by) of in, lower part only
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t r = c0 + ty + 8 * k, c = r0 + tx;
      tile[ty + 8 * k][tx] = (r < n && c < n && c <= r) ? in[r * ld + c] : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = r0 + ty + 8 * k, c = c0 + tx;
    if (r < n && c < n) out[r * ld + c] = (bx >= by) ? tile[tx][ty + 8 * k] : 0.f;
  }
}

// out [n x n] (pitch ldo) = in^T (pitch ldi)
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, int64_t ldi, float* __restrict__ out, int64_t ldo, int64_t n) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32, c0 = static_cast<int64_t>(blockIdx.x) * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = r0 + ty + 8 * k, c = c0 + tx;
    tile[ty + 8 * k][tx] = (r < n && c < n) ? in[r * ldi + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = c0 + ty + 8 * k, c = r0 + tx;
    if (r < n && c < n) out[r * ldo + c] = tile[tx][ty + 8 * k];
  }
}

// n x n block with pitch ld <- identity (whole block) or strict upper <- 0
template <int MODE>
__global__ void __launch_bounds__(256)
block_fill_kernel(float* __restrict__ A, int64_t n, int64_t ld) {
  const int64_t total = n * n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / n, c = i - r * n;
    if (MODE == 0) A[r * ld + c] = (r == c) ? 1.f : 0.f;
    else if (c > r) A[r * ld + c] = 0.f;
  }
}

int launch_ok(const char* what) {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ODF_OK : set_cuda_error(e, what);
}

// ---- the split GEMM on prepared slices ------------------------------------------------------------------------------
struct Operand {           // one prepared k-slice: hi / lo [rows x pitch] fp16, norms (unused by the linear tile), scale
  void *hi, *lo;
  float *sqn, *scale;
  int64_t rows_cap;
};

// ODF_PRECOND_TRACE=2: additionally the summed CUDA-event time per category of call inside the phases
struct CatTrace {
  bool on = false;
  struct Rec { int cat; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  cudaEvent_t open_ev = nullptr;
  int open_cat = -1;
  void begin(int cat, cudaStream_t st) {
    if (!on) return;
    cudaEventCreate(&open_ev);
    cudaEventRecord(open_ev, st);
    open_cat = cat;
  }
  void end(cudaStream_t st) {
    if (!on) return;
    cudaEvent_t b;
    cudaEventCreate(&b);
    cudaEventRecord(b, st);
    recs.push_back({open_cat, open_ev, b});
  }
  void report() {
    if (!on) return;
    static const char* names[] = {"potrf(diag)", "trsm(panel)", "prep", "tile(update)", "tile(ttt)", "tile(inverse)", "transpose"};
    float sum[7] = {0, 0, 0, 0, 0, 0, 0};
    int cnt[7] = {0, 0, 0, 0, 0, 0, 0};
    for (auto& r : recs) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.a, r.b);
      sum[r.cat] += ms;
      ++cnt[r.cat];
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    printf("precond_build categories:");
    for (int i = 0; i < 7; ++i) printf("  %s %.2f ms / %d", names[i], sum[i], cnt[i]);
    printf("\n");
    fflush(stdout);
  }
};
enum : int { CAT_POTRF = 0, CAT_TRSM, CAT_PREP, CAT_UPDATE, CAT_TTT, CAT_INV, CAT_TRANSPOSE };

struct Build {
  CatTrace cat;
  cudaStream_t st;
  cublasHandle_t blas;
  cusolverDnHandle_t sol;
  Operand a, b;
  float* scratch;          // [M x M]: X11^T of the node being combined
  float* diag;             // [nbc x nbc] contiguous copy of the diagonal block being factorised
  float* potrf_work;
  int lwork;
  int* info;               // device: one int per potrf call
  int n_info, info_cap;
  cudaStream_t side[N_SIDE];
  cublasHandle_t side_blas[N_SIDE];   // one handle per side stream: a handle's internal workspace is not stream-safe
  cudaEvent_t ev_fork, ev_join[N_SIDE];
};

int prep(Build& B, Operand& o, const float* X, int64_t ld, int64_t rows, int64_t k) {
  if (rows > o.rows_cap || k > KSLICE) return set_error(ODF_ERR_ARG, "precond_build: operand slice exceeds the workspace");
  B.cat.begin(CAT_PREP, B.st);
  const int rc = prepare_points(X, rows, k, ld, nullptr, 1.f, KIND_F16, o.hi, o.lo, o.sqn, o.scale, B.st, true);
  B.cat.end(B.st);
  return rc;
}

// C[m x n] (pitch ldc) = alpha * A_rows[a0 : a0 + m] . B_rows[b0 : b0 + n]^T + beta * C on already prepared slices of
// width k (pitch = operand_pitch(k)); a0, b0 multiples of 4.
int tile_nt(Build& B, const Operand& A, int64_t a0, int64_t m, const Operand& Bo, int64_t b0, int64_t n, int64_t k, float alpha,
            float beta, float* C, int64_t ldc, int cat) {
  if (m <= 0 || n <= 0) return ODF_OK;
  const int64_t pitch = operand_pitch(k, KIND_F16);
  TileLaunch L{};
  L.kind = KIND_F16;
  L.r_hi = static_cast<const __half*>(A.hi) + a0 * pitch;
  L.r_lo = static_cast<const __half*>(A.lo) + a0 * pitch;
  L.r_norm = A.sqn + a0; L.r_scale = A.scale; L.n_rows = m;
  L.q_hi = static_cast<const __half*>(Bo.hi) + b0 * pitch;
  L.q_lo = static_cast<const __half*>(Bo.lo) + b0 * pitch;
  L.q_norm = Bo.sqn + b0; L.q_scale = Bo.scale; L.n_cols = n;
  L.d_pad = round_up(k, kblock_elems(KIND_F16)); L.T_pad = 16; L.mode = MODE_STORE;
  L.n_splits = tile_default_splits(m, n, pitch * 2); L.sigma = 1.f;
  L.out = C; L.ldo = ldc; L.split_stride = 0;
  L.linear = 1; L.lin_alpha = alpha; L.lin_beta = beta;
  B.cat.begin(cat, B.st);
  const int rc = launch_gauss_tile(L, B.st);
  B.cat.end(B.st);
  return rc;
}

// ---- blocked Cholesky, row-major lower, in place -----------------------------------------------------------------------
int chol_lower(Build& B, float* G, int64_t M) {
  const float one = 1.f;
  const int64_t NBC = chol_block();
  int rc;
  for (int64_t j0 = 0; j0 < M; j0 += NBC) {
    const int64_t jb = (M - j0 < NBC) ? (M - j0) : NBC;
    float* D = G + j0 * M + j0;
    if (B.n_info >= B.info_cap) return set_error(ODF_ERR_ARG, "precond_build: too many diagonal blocks");
    B.cat.begin(CAT_POTRF, B.st);
    if (potrf_via_copy()) {
      // contiguous TRANSPOSED copy: only the row-major lower triangle of the block is valid (the updates write lower block
      // columns), and transposed it is the column-major lower triangle cuSOLVER's fast path reads (LOWER potrf: 0.35 ms at
      // 1024 against 0.92 ms for the strided UPPER call); the factor is written back transposed
      dim3 grid(static_cast<unsigned>((jb + 31) / 32), static_cast<unsigned>((jb + 31) / 32));
      transpose_kernel<<<grid, 256, 0, B.st>>>(D, M, B.diag, jb, jb);
      if ((rc = launch_ok("transpose (diagonal block in)"))) return rc;
      if (cusolverDnSpotrf(B.sol, CUBLAS_FILL_MODE_LOWER, static_cast<int>(jb), B.diag, static_cast<int>(jb), B.potrf_work, B.lwork,
                           B.info + B.n_info++) != CUSOLVER_STATUS_SUCCESS)
        return set_error(ODF_ERR_CUDA, "cusolverDnSpotrf (diagonal block) failed to launch");
      transpose_kernel<<<grid, 256, 0, B.st>>>(B.diag, jb, D, M, jb);      // row-major lower L11 <- (column-major lower)^T
      if ((rc = launch_ok("transpose (diagonal block)"))) return rc;
    } else {
      // row-major lower L11 == column-major upper U = L11^T
      if (cusolverDnSpotrf(B.sol, CUBLAS_FILL_MODE_UPPER, static_cast<int>(jb), D, static_cast<int>(M), B.potrf_work, B.lwork,
                           B.info + B.n_info++) != CUSOLVER_STATUS_SUCCESS)
        return set_error(ODF_ERR_CUDA, "cusolverDnSpotrf (diagonal block) failed to launch");
    }
    B.cat.end(B.st);
    const int64_t m = M - j0 - jb;
    if (m == 0) break;
    float* P = G + (j0 + jb) * M + j0;                     // row panel [m x jb], pitch M
    // L21 = G21 L11^-T: in column-major terms P' = L21^T (jb x m) solves U^T P' = G21^T
    B.cat.begin(CAT_TRSM, B.st);
    if (cublasStrsm(B.blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, static_cast<int>(jb),
                    static_cast<int>(m), &one, D, static_cast<int>(M), P, static_cast<int>(M)) != CUBLAS_STATUS_SUCCESS)
      return set_error(ODF_ERR_CUDA, "cublasStrsm (Cholesky panel) failed");
    B.cat.end(B.st);
    // trailing update, lower block columns only, one k-slice of the panel at a time:
    //   G[J0:M, J0:J1] -= L[J0:M, slice] L[J0:J1, slice]^T
    for (int64_t ks = 0; ks < jb; ks += KSLICE) {
      const int64_t kw = (jb - ks < KSLICE) ? (jb - ks) : KSLICE;
      if ((rc = prep(B, B.a, P + ks, M, m, kw))) return rc;
      for (int64_t J0 = j0 + jb; J0 < M; J0 += NB) {
        const int64_t Jb = (M - J0 < NB) ? (M - J0) : NB;
        const int64_t off = J0 - (j0 + jb);
        if ((rc = tile_nt(B, B.a, off, M - J0, B.a, off, Jb, kw, -1.f, 1.f, G + J0 * M + J0, M, CAT_UPDATE))) return rc;
      }
    }
  }
  return ODF_OK;
}

// ---- lower block of T T^T / M for an UPPER-triangular T (row-major), G pre-zeroed ----------------------------------------
int ttt_lower(Build& B, const float* T, float* G, int64_t M) {
  int rc;
  const float inv_m = 1.f / static_cast<float>(M);
  for (int64_t k0 = 0; k0 < M; k0 += KSLICE) {
    const int64_t k1 = (M - k0 < KSLICE) ? M : k0 + KSLICE;
    // rows >= k1 of this column slice are zero (T upper): only T[0:k1, k0:k1] contributes
    if ((rc = prep(B, B.a, T + k0, M, k1, k1 - k0))) return rc;
    for (int64_t J0 = 0; J0 < k1; J0 += NB) {
      const int64_t Jb = (k1 - J0 < NB) ? (k1 - J0) : NB;
      if ((rc = tile_nt(B, B.a, J0, k1 - J0, B.a, J0, Jb, k1 - k0, inv_m, 1.f, G + J0 * M + J0, M, CAT_TTT))) return rc;
    }
  }
  return ODF_OK;
}

// ---- X = L^-1 (both row-major lower, pitch ld), X zero outside the leaves on entry ---------------------------------------
int64_t split_point(int64_t n) { return round_up((n + 1) / 2, NB); }

// leaves first, round-robin on the side streams (independent, each far too small to fill the GPU)
int inv_leaves(Build& B, const float* L, float* X, int64_t n, int64_t ld, int* counter) {
  if (n <= NB) {
    const int slot = (*counter)++ % N_SIDE;
    cudaStream_t s = B.side[slot];
    const float one = 1.f;
    block_fill_kernel<0><<<static_cast<unsigned>((n * n + 255) / 256 > 1184 ? 1184 : (n * n + 255) / 256), 256, 0, s>>>(X, n, ld);
    // column-major: L is the upper U = L^T; U Y = I gives Y = U^-1 = (L^-1)^T, i.e. row-major X = L^-1
    cublasStatus_t cs = cublasStrsm(B.side_blas[slot], CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT,
                                    static_cast<int>(n), static_cast<int>(n), &one, L, static_cast<int>(ld), X, static_cast<int>(ld));
    if (cs != CUBLAS_STATUS_SUCCESS) return set_error(ODF_ERR_CUDA, "cublasStrsm (leaf inverse) failed");
    block_fill_kernel<1><<<static_cast<unsigned>((n * n + 255) / 256 > 1184 ? 1184 : (n * n + 255) / 256), 256, 0, s>>>(X, n, ld);
    return launch_ok("leaf inverse");
  }
  const int64_t n1 = split_point(n);
  int rc;
  if ((rc = inv_leaves(B, L, X, n1, ld, counter))) return rc;
  return inv_leaves(B, L + n1 * ld + n1, X + n1 * ld + n1, n - n1, ld, counter);
}

int inv_combine(Build& B, const float* L, float* X, int64_t n, int64_t ld) {
  if (n <= NB) return ODF_OK;
  const int64_t n1 = split_point(n), n2 = n - n1;
  int rc;
  if ((rc = inv_combine(B, L, X, n1, ld))) return rc;
  if ((rc = inv_combine(B, L + n1 * ld + n1, X + n1 * ld + n1, n2, ld))) return rc;
  const float* L21 = L + n1 * ld;            // [n2 x n1]
  float* X12 = X + n1;                        // [n1 x n2], zero: scratch for W^T = X11^T L21^T
  float* X21 = X + n1 * ld;                   // [n2 x n1], zero on entry
  const float* X22 = X + n1 * ld + n1;
  // X11^T (upper) into the scratch matrix
  {
    dim3 grid(static_cast<unsigned>((n1 + 31) / 32), static_cast<unsigned>((n1 + 31) / 32));
    B.cat.begin(CAT_TRANSPOSE, B.st);
    transpose_kernel<<<grid, 256, 0, B.st>>>(X, ld, B.scratch, n1, n1);
    B.cat.end(B.st);
    if ((rc = launch_ok("transpose (X11)"))) return rc;
  }
  // W^T[0:k1, :] += X11^T[0:k1, k0:k1] . L21[:, k0:k1]^T   (X11^T upper: rows >= k1 of the slice are zero)
  for (int64_t k0 = 0; k0 < n1; k0 += KSLICE) {
    const int64_t k1 = (n1 - k0 < KSLICE) ? n1 : k0 + KSLICE;
    if ((rc = prep(B, B.a, B.scratch + k0, n1, k1, k1 - k0))) return rc;
    if ((rc = prep(B, B.b, L21 + k0, ld, n2, k1 - k0))) return rc;
    if ((rc = tile_nt(B, B.a, 0, k1, B.b, 0, n2, k1 - k0, 1.f, 1.f, X12, ld, CAT_INV))) return rc;
  }
  // X21[k0:n2, :] -= X22[k0:n2, k0:k1] . W^T[:, k0:k1]^T    (X22 lower: rows < k0 of the slice are zero)
  for (int64_t k0 = 0; k0 < n2; k0 += KSLICE) {
    const int64_t k1 = (n2 - k0 < KSLICE) ? n2 : k0 + KSLICE;
    if ((rc = prep(B, B.a, X22 + k0 * ld + k0, ld, n2 - k0, k1 - k0))) return rc;
    if ((rc = prep(B, B.b, X12 + k0, ld, n1, k1 - k0))) return rc;
    if ((rc = tile_nt(B, B.a, 0, n2 - k0, B.b, 0, n1, k1 - k0, -1.f, 1.f, X21 + k0 * ld, ld, CAT_INV))) return rc;
  }
  cudaError_t e = cudaMemset2DAsync(X12, static_cast<size_t>(ld) * 4, 0, static_cast<size_t>(n2) * 4, static_cast<size_t>(n1), B.st);
  return e == cudaSuccess ? ODF_OK : set_cuda_error(e, "precond_build: memset (scratch block)");
}

int inv_lower(Build& B, const float* L, float* X, int64_t M) {
  cudaError_t e = cudaMemsetAsync(X, 0, static_cast<size_t>(M) * M * 4, B.st);
  if (e != cudaSuccess) return set_cuda_error(e, "precond_build: memset (inverse)");
  // fork: the leaves run on the side streams behind everything queued so far, the combination waits for all of them
  if ((e = cudaEventRecord(B.ev_fork, B.st)) != cudaSuccess) return set_cuda_error(e, "precond_build: event");
  for (int s = 0; s < N_SIDE; ++s)
    if ((e = cudaStreamWaitEvent(B.side[s], B.ev_fork, 0)) != cudaSuccess) return set_cuda_error(e, "precond_build: fork");
  int counter = 0, rc;
  if ((rc = inv_leaves(B, L, X, M, M, &counter))) return rc;
  for (int s = 0; s < N_SIDE; ++s) {
    if ((e = cudaEventRecord(B.ev_join[s], B.side[s])) != cudaSuccess) return set_cuda_error(e, "precond_build: event");
    if ((e = cudaStreamWaitEvent(B.st, B.ev_join[s], 0)) != cudaSuccess) return set_cuda_error(e, "precond_build: join");
  }
  return inv_combine(B, L, X, M, M);
}

int to_upper(Build& B, float* A, int64_t M) {
  const int64_t nt = (M + 31) / 32;
  tril_to_triu_kernel<<<static_cast<unsigned>(nt * (nt + 1) / 2), 256, 0, B.st>>>(A, M, M, static_cast<int>(nt));
  return launch_ok("tril_to_triu_kernel");
}

// ODF_PRECOND_TRACE=1: CUDA-event time of every phase of the build, printed after the final synchronisation
struct Trace {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> names;
  cudaStream_t st = nullptr;
  void start(cudaStream_t s) {
    const char* e = getenv("ODF_PRECOND_TRACE");
    on = e && atoi(e) != 0;
    st = s;
    if (on) mark(nullptr);
  }
  void mark(const char* name) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    if (name) names.push_back(name);
  }
  void report(int64_t M) {
    if (!on) return;
    printf("precond_build M=%lld:", static_cast<long long>(M));
    for (size_t i = 0; i + 1 < ev.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      printf("  %s %.2f", names[i], ms);
    }
    printf("  (ms)\n");
    fflush(stdout);
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  }
};

// per-device side streams / events, created once: two sets (the main build and the inverse of T that runs beside it)
struct SideStreams {
  bool made = false;
  cudaStream_t side[N_SIDE];
  cublasHandle_t blas[N_SIDE];
  cudaStream_t aux;
  cudaEvent_t fork[2], join[2][N_SIDE], aux_go, aux_done;
};
SideStreams g_side[kMaxDevices];

int side_streams(Build& B, int set, cudaStream_t* aux, cudaEvent_t* aux_go, cudaEvent_t* aux_done) {
  SideStreams& S = g_side[current_device()];
  if (!S.made) {
    bool ok = cudaStreamCreateWithFlags(&S.aux, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&S.aux_go, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&S.aux_done, cudaEventDisableTiming) == cudaSuccess;
    for (int s = 0; ok && s < N_SIDE; ++s) {
      ok = cudaStreamCreateWithFlags(&S.side[s], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&S.join[0][s], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&S.join[1][s], cudaEventDisableTiming) == cudaSuccess &&
           cublasCreate(&S.blas[s]) == CUBLAS_STATUS_SUCCESS;
      if (ok) {
        cublasSetMathMode(S.blas[s], CUBLAS_DEFAULT_MATH);          // true fp32
        cublasSetStream(S.blas[s], S.side[s]);
      }
    }
    ok = ok && cudaEventCreateWithFlags(&S.fork[0], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&S.fork[1], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) return set_error(ODF_ERR_CUDA, "precond_build: cannot create side streams");
    S.made = true;
  }
  for (int s = 0; s < N_SIDE; ++s) { B.side[s] = S.side[s]; B.ev_join[s] = S.join[set][s]; B.side_blas[s] = S.blas[s]; }
  B.ev_fork = S.fork[set];
  if (aux) { *aux = S.aux; *aux_go = S.aux_go; *aux_done = S.aux_done; }
  return ODF_OK;
}

size_t operand_slice_bytes(int64_t rows) { return al256(static_cast<size_t>(rows) * operand_pitch(KSLICE, KIND_F16) * 2); }
size_t operand_set_bytes(int64_t M) { return 2 * operand_slice_bytes(M) + al256(sizeof(float) * round_up(M, 128)) + 256; }

}  // namespace

size_t precond_build_workspace_bytes(int64_t M, int lwork) {
  const int64_t n_blocks = (M + NB - 1) / NB;
  const int64_t nbc = M < NBC_MAX ? M : NBC_MAX;
  // two builds' worth of operand slices and scratch (the inverse of T runs on its own stream beside T T^T / chol(A))
  return 4 * operand_set_bytes(M) + 2 * al256(static_cast<size_t>(M) * M * 4) + al256(static_cast<size_t>(nbc) * nbc * 4) +
         al256(sizeof(float) * static_cast<size_t>(lwork > 0 ? lwork : 1)) + al256(sizeof(int) * static_cast<size_t>(2 * n_blocks + 2)) + 1024;
}

// K: K_MM in, destroyed (it ends up holding the lower factor of T).  Tm, Am: T, A (upper).  Tinv, Ainv: their inverses or NULL.
int precond_build(float* K, float* Tm, float* Am, float* Tinv, float* Ainv, int64_t M, float lam, float eps, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
  if (M <= 0 || M > 0x7fffffff || !K || !Tm || !Am || K == Tm) return set_error(ODF_ERR_ARG, "precond_build: bad arguments");
  if ((Tinv == nullptr) != (Ainv == nullptr)) return set_error(ODF_ERR_ARG, "precond_build: Tinv and Ainv go together");
  Build B{}, B2{};
  B.st = st;
  int rc;
  if ((rc = lib_handles(st, &B.blas, &B.sol))) return rc;
  cudaStream_t aux;
  cudaEvent_t aux_go, aux_done;
  if ((rc = side_streams(B, 0, &aux, &aux_go, &aux_done))) return rc;
  if ((rc = side_streams(B2, 1, nullptr, nullptr, nullptr))) return rc;
  B2.st = aux;
  const int nbq = static_cast<int>(M < chol_block() ? M : chol_block());
  if (cusolverDnSpotrf_bufferSize(B.sol, CUBLAS_FILL_MODE_LOWER, nbq, K, nbq, &B.lwork) != CUSOLVER_STATUS_SUCCESS)
    return set_error(ODF_ERR_CUDA, "cusolverDnSpotrf_bufferSize failed");
  int lw2 = 0;
  if (cusolverDnSpotrf_bufferSize(B.sol, CUBLAS_FILL_MODE_UPPER, nbq, K, static_cast<int>(M), &lw2) == CUSOLVER_STATUS_SUCCESS && lw2 > B.lwork)
    B.lwork = lw2;
  if (ws_bytes < precond_build_workspace_bytes(M, B.lwork)) return set_error(ODF_ERR_WORKSPACE, "precond_build: workspace too small");
  // carve the workspace
  uint8_t* p = static_cast<uint8_t*>(ws);
  auto take = [&](size_t bytes) { uint8_t* r = p; p += al256(bytes); return r; };
  for (Operand* o : {&B.a, &B.b, &B2.a, &B2.b}) {
    o->hi = take(operand_slice_bytes(M));
    o->lo = take(operand_slice_bytes(M));
    o->sqn = reinterpret_cast<float*>(take(sizeof(float) * round_up(M, 128)));
    o->scale = reinterpret_cast<float*>(take(256));
    o->rows_cap = M;
  }
  B.scratch = reinterpret_cast<float*>(take(static_cast<size_t>(M) * M * 4));
  B2.scratch = reinterpret_cast<float*>(take(static_cast<size_t>(M) * M * 4));
  B.diag = reinterpret_cast<float*>(take(static_cast<size_t>(nbq) * nbq * 4));
  B.potrf_work = reinterpret_cast<float*>(take(sizeof(float) * static_cast<size_t>(B.lwork > 0 ? B.lwork : 1)));
  B.info_cap = static_cast<int>(2 * ((M + NB - 1) / NB) + 2);
  B.info = reinterpret_cast<int*>(take(sizeof(int) * static_cast<size_t>(B.info_cap)));
  B.n_info = 0;
  cudaError_t e = cudaMemsetAsync(B.info, 0, sizeof(int) * static_cast<size_t>(B.info_cap), st);
  if (e != cudaSuccess) return set_cuda_error(e, "precond_build: memset (info)");

  Trace tr;
  tr.start(st);
  {
    const char* e2 = getenv("ODF_PRECOND_TRACE");
    B.cat.on = e2 && atoi(e2) >= 2;
  }
  const dim3 tgrid(static_cast<unsigned>((M + 31) / 32), static_cast<unsigned>((M + 31) / 32));
  // T: K_MM + eps M I = L L^T (in place in K)
  if ((rc = add_diag(K, M, eps * static_cast<float>(M), st))) return rc;
  if ((rc = chol_lower(B, K, M))) return rc;
  const int n_info_T = B.n_info;
  tr.mark("chol(T)");
  // X = L^-1 on the aux stream, beside T T^T and the second factorisation (which are latency-bound on the diagonal blocks)
  if (Tinv) {
    if ((e = cudaEventRecord(aux_go, st)) != cudaSuccess || (e = cudaStreamWaitEvent(aux, aux_go, 0)) != cudaSuccess)
      return set_cuda_error(e, "precond_build: fork (aux)");
    if ((rc = inv_lower(B2, K, Tinv, M))) return rc;
    if ((rc = to_upper(B2, Tinv, M))) return rc;
    if ((e = cudaEventRecord(aux_done, aux)) != cudaSuccess) return set_cuda_error(e, "precond_build: event (aux)");
  }
  tril_transpose_copy_kernel<<<tgrid, 256, 0, st>>>(K, Tm, M, M);                  // T = L^T, upper, strict lower zeroed
  if ((rc = launch_ok("tril_transpose_copy_kernel"))) return rc;
  // A: T T^T / M + lam I = L_A L_A^T
  if ((e = cudaMemsetAsync(Am, 0, static_cast<size_t>(M) * M * 4, st)) != cudaSuccess) return set_cuda_error(e, "precond_build: memset (A)");
  if ((rc = ttt_lower(B, Tm, Am, M))) return rc;
  if ((rc = add_diag(Am, M, lam, st))) return rc;
  tr.mark("T^T copy + T T^T");
  if ((rc = chol_lower(B, Am, M))) return rc;
  tr.mark("chol(A)");
  if (Ainv && (rc = inv_lower(B, Am, Ainv, M))) return rc;
  tr.mark("inv(A)");
  if ((rc = to_upper(B, Am, M))) return rc;
  if (Ainv && (rc = to_upper(B, Ainv, M))) return rc;
  if (Tinv && (e = cudaStreamWaitEvent(st, aux_done, 0)) != cudaSuccess) return set_cuda_error(e, "precond_build: join (aux)");
  tr.mark("transposes + join inv(T)");

  std::vector<int> hinfo(static_cast<size_t>(B.n_info), 0);
  e = cudaMemcpyAsync(hinfo.data(), B.info, sizeof(int) * hinfo.size(), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_cuda_error(e, "precond_build");
  tr.report(M);
  B.cat.report();
  for (size_t i = 0; i < hinfo.size(); ++i) {
    if (hinfo[i] != 0) {
      char buf[200];
      const bool in_T = static_cast<int>(i) < n_info_T;
      snprintf(buf, sizeof buf, "Cholesky failed in the %s factor: diagonal block %d, info=%d (matrix not positive definite)",
               in_T ? "T" : "A", static_cast<int>(i) - (in_T ? 0 : n_info_T), hinfo[i]);
      return set_error(ODF_ERR_LINALG, buf);
    }
  }
  return ODF_OK;
}

}  // namespace odf
