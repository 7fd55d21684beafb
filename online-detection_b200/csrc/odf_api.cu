// extern "C" surface of libodf (declared in include/odf.h) + the cuSOLVER/cuBLAS-backed
// preconditioner.  Everything here is thin: argument checks, workspace carving, launches.
#include <cstring>
#include <string>

#include <cublas_v2.h>
#include <cusolverDn.h>

#include "odf_internal.h"

namespace odf {

// ---- implemented in odf_vec.cu
int prepare_points(const float*, int64_t, int64_t, int64_t, const float*, float, int, void*, void*, float*, float*, cudaStream_t,
                   bool zero_seed = false);
int zscore(float*, int64_t, int64_t, int64_t, const float*, float, cudaStream_t);
int split_rhs(const float*, int64_t, int64_t, int64_t, float, float*, float*, int64_t, int, cudaStream_t);
int finish_rows(const float*, int, int64_t, int, int64_t, float, const float*, int64_t, float*, int64_t, cudaStream_t);
int finish_split(const float*, int, int64_t, int, int64_t, float, const float*, int64_t, float*, float*, int64_t, cudaStream_t);
size_t cg_workspace_bytes(int64_t, int64_t);
int cg_init(const float*, int64_t, int64_t, int64_t, float*, void*, size_t, cudaStream_t);
int cg_alpha(const float*, const float*, int64_t, int64_t, int64_t, float, float*, void*, size_t, cudaStream_t);
int cg_beta(const float*, int64_t, int64_t, int64_t, float, float, float*, void*, size_t, cudaStream_t);
int cg_axpy_a(float*, const float*, int64_t, int64_t, int64_t, float, const float*, cudaStream_t);
int cg_xpby_b(float*, const float*, int64_t, int64_t, int64_t, const float*, cudaStream_t);
int cg_residual(float*, const float*, const float*, int64_t, int64_t, int64_t, const float*, cudaStream_t);
int axpby(float*, float, const float*, float, const float*, int64_t, int64_t, int64_t, cudaStream_t);
int add_diag(float*, int64_t, float, cudaStream_t);
int zero_lower(float*, int64_t, cudaStream_t);
int set_identity(float*, int64_t, cudaStream_t);

namespace {
thread_local std::string g_err = "";
}

int set_error(int code, const char* msg) {
  g_err = msg ? msg : "";
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  g_err = std::string(where ? where : "cuda") + ": " + cudaGetErrorString(e);
  return ODF_ERR_CUDA;
}

namespace {

// bump allocator over the caller's workspace (256-byte granules)
struct Arena {
  uint8_t* base;
  size_t cap, off;
  Arena(void* p, size_t n) : base(static_cast<uint8_t*>(p)), cap(n), off(0) {}
  template <class T>
  T* take(size_t count) {
    const size_t bytes = (count * sizeof(T) + 255) & ~static_cast<size_t>(255);
    if (base == nullptr || off + bytes > cap) { off = cap + 1; return nullptr; }
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= cap; }
};
inline size_t al(size_t bytes) { return (bytes + 255) & ~static_cast<size_t>(255); }

int g_default_kind = ODF_KIND_F16;

struct Prepared {
  void *hi, *lo;
  float *sqn, *opscale;
};
size_t operand_bytes(int64_t n, int64_t d, int kind) {
  return static_cast<size_t>(n) * operand_pitch(d, kind) * (kind == KIND_F16 ? 2 : 4);
}
size_t prepared_bytes(int64_t n, int64_t d, int kind) {
  return 2 * al(operand_bytes(n, d, kind)) + al(sizeof(float) * round_up(n, 128)) + 256;
}
bool take_prepared(Arena& a, int64_t n, int64_t d, int kind, Prepared* p) {
  p->hi = a.take<uint8_t>(operand_bytes(n, d, kind));
  p->lo = a.take<uint8_t>(operand_bytes(n, d, kind));
  p->sqn = a.take<float>(round_up(n, 128));
  p->opscale = a.take<float>(2);
  return a.ok();
}

int tpad_of(int64_t T) { return T <= 16 ? 16 : (T <= 32 ? 32 : -1); }

// library handles, one pair per device (a handle is bound to the device that was current when it was created);
// the references below are re-pointed at the current device's pair by ensure_handles().  One host thread per device.
cublasHandle_t g_cublas_dev[kMaxDevices] = {};
cusolverDnHandle_t g_cusolver_dev[kMaxDevices] = {};
thread_local cublasHandle_t g_cublas = nullptr;
thread_local cusolverDnHandle_t g_cusolver = nullptr;
int ensure_handles(cudaStream_t st) {
  const int dev = current_device();
  g_cublas = g_cublas_dev[dev];
  g_cusolver = g_cusolver_dev[dev];
  struct Publish {                 // store freshly created handles back on every exit path
    int dev;
    ~Publish() { g_cublas_dev[dev] = g_cublas; g_cusolver_dev[dev] = g_cusolver; }
  } publish{dev};
  if (!g_cublas) {
    if (cublasCreate(&g_cublas) != CUBLAS_STATUS_SUCCESS) return set_error(ODF_ERR_CUDA, "cublasCreate failed");
    // triangular solves / syrk of the preconditioner stay in true fp32
    cublasSetMathMode(g_cublas, CUBLAS_DEFAULT_MATH);  // fp32 routines stay true fp32 (no TF32)
  }
  if (!g_cusolver) {
    if (cusolverDnCreate(&g_cusolver) != CUSOLVER_STATUS_SUCCESS) return set_error(ODF_ERR_CUDA, "cusolverDnCreate failed");
  }
  cublasSetStream(g_cublas, st);
  cusolverDnSetStream(g_cusolver, st);
  return ODF_OK;
}

// Row-major C (m x n) = alpha op(A) op(B) + beta C, true fp32 (cuBLAS SIMT sgemm, ~67 TFLOP/s on B200).
// cuBLAS 12.9 can emulate this GEMM on the bf16 tensor cores (CUBLAS_COMPUTE_32F_EMULATED_16BFX9: 175-200 TFLOP/s with
// better-than-sgemm accuracy, tools/emu_probe.cu), but inside a PyTorch process the cuBLAS that is already loaded is
// torch's own 12.8 build, which rejects that compute type -- so the preconditioner gets its speed from doing less work
// (triangle-aware products, recursive inverse) instead.  ODF_PRECOND_TRACE=1 prints every call with its timing.
int gemm_rm(bool ta, bool tb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda, const float* B,
            int64_t ldb, float beta, float* C, int64_t ldc) {
  static int trace = -1;
  if (trace < 0) {
    const char* e = getenv("ODF_PRECOND_TRACE");
    trace = (e && atoi(e) != 0) ? 1 : 0;
  }
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t cst = nullptr;
  if (trace) {
    cublasGetStream(g_cublas, &cst);
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    cudaEventRecord(ev0, cst);
  }
  // row-major C = op(A) op(B)  <=>  column-major C' = op(B') op(A')
  cublasStatus_t st = cublasSgemm(g_cublas, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, static_cast<int>(n),
                                  static_cast<int>(m), static_cast<int>(k), &alpha, B, static_cast<int>(ldb), A,
                                  static_cast<int>(lda), &beta, C, static_cast<int>(ldc));
  if (trace) {
    cudaEventRecord(ev1, cst);
    cudaEventSynchronize(ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    printf("gemm_rm %c%c m=%lld n=%lld k=%lld: %.3f ms  %.1f TFLOP/s  status=%d\n", ta ? 'T' : 'N', tb ? 'T' : 'N', (long long)m,
           (long long)n, (long long)k, ms, 2.0 * m * n * k / ms / 1e9, (int)st);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
  }
  return st == CUBLAS_STATUS_SUCCESS ? ODF_OK : set_error(ODF_ERR_CUDA, "cublasSgemm (preconditioner) failed");
}

constexpr int64_t PC_NB = 2048;     // block width of the triangular-aware products
constexpr int64_t PC_LEAF = 1024;   // diagonal blocks inverted by one TRSM against the identity

// Upper triangle (row-major, block granularity) of A = alpha T T^T for an upper-triangular T (zeros below the diagonal):
// A[0:(J+1)nb, Jblock] = T[0:(J+1)nb, J nb:] . T[Jblock, J nb:]^T  -- M^3/3 useful flops instead of the 2 M^3 of a full GEMM.
int ttt_upper(const float* T, float* A, int64_t M, float alpha) {
  for (int64_t c0 = 0; c0 < M; c0 += PC_NB) {
    const int64_t nb = (M - c0 < PC_NB) ? (M - c0) : PC_NB;
    const int rc = gemm_rm(false, true, c0 + nb, nb, M - c0, alpha, T + c0, M, T + c0 * M + c0, M, 0.f, A + c0, M);
    if (rc) return rc;
  }
  return ODF_OK;
}

// X = T^-1 for an upper-triangular n x n block (row-major, pitch ld).  On entry X holds the identity.  Recursive:
//   [T11 T12; 0 T22]^-1 = [X11, -X11 T12 X22; 0, X22];  the (otherwise zero) X21 block is the scratch for (T12 X22)^T.
int inv_upper_rec(const float* T, float* X, int64_t n, int64_t ld, cudaStream_t st) {
  if (n <= PC_LEAF) {
    // X' L = I' with L = T^T column-major lower (see odf_precond_solve)
    const float one = 1.f;
    if (cublasStrsm(g_cublas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, static_cast<int>(n),
                    static_cast<int>(n), &one, T, static_cast<int>(ld), X, static_cast<int>(ld)) != CUBLAS_STATUS_SUCCESS)
      return set_error(ODF_ERR_CUDA, "cublasStrsm (precond_invert leaf) failed");
    return ODF_OK;
  }
  const int64_t n1 = ((n / 2) + 127) / 128 * 128, n2 = n - n1;
  const float* T12 = T + n1;
  const float* T22 = T + n1 * ld + n1;
  float* X12 = X + n1;
  float* X21 = X + n1 * ld;
  float* X22 = X + n1 * ld + n1;
  int rc;
  if ((rc = inv_upper_rec(T, X, n1, ld, st))) return rc;
  if ((rc = inv_upper_rec(T22, X22, n2, ld, st))) return rc;
  if ((rc = gemm_rm(true, true, n2, n1, n2, 1.f, X22, ld, T12, ld, 0.f, X21, ld))) return rc;     // X21 <- (T12 X22)^T
  if ((rc = gemm_rm(false, true, n1, n2, n1, -1.f, X, ld, X21, ld, 0.f, X12, ld))) return rc;      // X12 <- -X11 (T12 X22)
  // the scratch block is part of the (upper-triangular) result of the caller's level: back to exact zeros
  cudaError_t e = cudaMemset2DAsync(X21, static_cast<size_t>(ld) * 4, 0, static_cast<size_t>(n1) * 4, static_cast<size_t>(n2), st);
  return e == cudaSuccess ? ODF_OK : set_cuda_error(e, "precond_invert: memset");
}

}  // namespace

// the current device's library handles, bound to `st` (odf_precond.cu)
int lib_handles(cudaStream_t st, cublasHandle_t* blas, cusolverDnHandle_t* solver) {
  const int rc = ensure_handles(st);
  if (rc) return rc;
  *blas = g_cublas;
  *solver = g_cusolver;
  return ODF_OK;
}
int tri_apply(const float* U, int64_t M, const float* Bm, int64_t ldb, int64_t T, float* out, int64_t ldo, int64_t row0, int64_t row1,
              int transposed, cudaStream_t st);
size_t rls_workspace_bytes(int64_t n, int64_t d, int64_t n_classes, int lwork);
int rls_query_lwork(int64_t d, int* lwork);
int rls_train(const float* X, int64_t n, int64_t d, int64_t ldx, const double* Yw, const int64_t* perm, const int64_t* seg_host,
              const int* row_class, int64_t n_classes, double lam, float* W, float* losses, void* ws, size_t ws_bytes,
              cudaStream_t st);
int rls_apply(const float* feat, int64_t n, int64_t d, int64_t ldf, const float* Wp, const float* bias, const float* Tinv,
              const float* mu, const float* ex_boxes, int64_t C, float img_w, float img_h, float eps, const float* mean, float zscale,
              float* out, cudaStream_t st);
size_t precond_build_workspace_bytes(int64_t M, int lwork);
int precond_build(float* K, float* Tm, float* Am, float* Tinv, float* Ainv, int64_t M, float lam, float eps, void* ws,
                  size_t ws_bytes, cudaStream_t st);
}  // namespace odf

using namespace odf;

extern "C" {

const char* odf_last_error(void) { return g_err.c_str(); }
int odf_version(void) { return 100; }

int odf_set_default_kind(int kind) {
  if (kind != ODF_KIND_TF32 && kind != ODF_KIND_F16) return set_error(ODF_ERR_ARG, "unknown operand kind");
  g_default_kind = kind;
  return ODF_OK;
}
int odf_default_kind(void) { return g_default_kind; }

int64_t odf_operand_pitch(int64_t d, int kind) { return operand_pitch(d, kind); }
int64_t odf_operand_bytes(int64_t n, int64_t d, int kind) { return static_cast<int64_t>(operand_bytes(n, d, kind)); }
int64_t odf_pad_rows(int64_t n) { return round_up(n, 128); }
int odf_tpad(int64_t T) { return tpad_of(T); }
int odf_tile_splits(int64_t n_rows, int64_t n_cols, int64_t d, int kind) {
  return tile_default_splits(n_rows, n_cols, operand_pitch(d, kind) * (kind == KIND_F16 ? 2 : 4));
}

int odf_prepare_points(const float* X, int64_t n, int64_t d, int64_t ldx, const float* mean, float scale,
                       int kind, void* hi, void* lo, float* sqnorm, float* opscale, void* stream) {
  return prepare_points(X, n, d, ldx, mean, scale, kind, hi, lo, sqnorm, opscale, static_cast<cudaStream_t>(stream));
}
int odf_prepare_points_linear(const float* X, int64_t n, int64_t d, int64_t ldx, int kind, void* hi, void* lo,
                              float* sqnorm, float* opscale, void* stream) {
  return prepare_points(X, n, d, ldx, nullptr, 1.f, kind, hi, lo, sqnorm, opscale, static_cast<cudaStream_t>(stream), true);
}
int odf_zscore(float* X, int64_t n, int64_t d, int64_t ldx, const float* mean, float scale, void* stream) {
  return zscore(X, n, d, ldx, mean, scale, static_cast<cudaStream_t>(stream));
}
int odf_split_rhs(const float* V, int64_t m, int64_t T, int64_t ldv, float scale, float* vt_hi, float* vt_lo,
                  int64_t ldvt, int T_pad, void* stream) {
  return split_rhs(V, m, T, ldv, scale, vt_hi, vt_lo, ldvt, T_pad, static_cast<cudaStream_t>(stream));
}

int odf_gauss_mmv_prepared(int kind, const void* r_hi, const void* r_lo, const float* r_sqnorm,
                           const float* r_opscale, int64_t n_rows, const void* q_hi, const void* q_lo,
                           const float* q_sqnorm, const float* q_opscale, int64_t n_cols, int64_t d,
                           const float* vt_hi, const float* vt_lo, int64_t ldvt, int T_pad, int n_splits,
                           float sigma, float* partial, void* stream) {
  return odf_gauss_mmv_prepared_spill(kind, r_hi, r_lo, r_sqnorm, r_opscale, n_rows, q_hi, q_lo, q_sqnorm, q_opscale,
                                      n_cols, d, vt_hi, vt_lo, ldvt, T_pad, n_splits, sigma, partial, nullptr, 0, stream);
}

int odf_panel_splits(int64_t n_rows, int64_t M) { return panel_splits(n_rows, M); }
int odf_panel_tmm(const float* P, int64_t ldp, const float* W, int64_t n_rows, int64_t M, int T_pad, int n_splits,
                  float* out_partial, void* stream) {
  return launch_panel_tmm(P, ldp, W, n_rows, M, T_pad, n_splits, out_partial, static_cast<cudaStream_t>(stream));
}

int odf_gauss_mmv_prepared_spill(int kind, const void* r_hi, const void* r_lo, const float* r_sqnorm,
                                 const float* r_opscale, int64_t n_rows, const void* q_hi, const void* q_lo,
                                 const float* q_sqnorm, const float* q_opscale, int64_t n_cols, int64_t d,
                                 const float* vt_hi, const float* vt_lo, int64_t ldvt, int T_pad, int n_splits,
                                 float sigma, float* partial, float* panel, int64_t ldp, void* stream) {
  if (!(sigma > 0.f)) return set_error(ODF_ERR_ARG, "sigma must be positive");
  if (kind != KIND_TF32 && kind != KIND_F16) return set_error(ODF_ERR_ARG, "unknown operand kind");
  TileLaunch L{};
  L.kind = kind;
  L.r_hi = r_hi; L.r_lo = r_lo; L.r_norm = r_sqnorm; L.r_scale = r_opscale; L.n_rows = n_rows;
  L.q_hi = q_hi; L.q_lo = q_lo; L.q_norm = q_sqnorm; L.q_scale = q_opscale; L.n_cols = n_cols;
  L.d_pad = round_up(d, kblock_elems(kind)); L.vt_hi = vt_hi; L.vt_lo = vt_lo; L.ldvt = ldvt; L.T_pad = T_pad;
  L.mode = MODE_MMV; L.n_splits = n_splits; L.sigma = sigma;
  L.out = partial; L.ldo = T_pad; L.split_stride = n_rows * T_pad;
  L.panel = panel; L.ldpanel = ldp;
  return launch_gauss_tile(L, static_cast<cudaStream_t>(stream));
}

int odf_tile_pair_eligible(int64_t n_rows) { return tile2_rows_eligible(n_rows) ? 1 : 0; }
int odf_split_rhs16(const float* V, int64_t m, int64_t T, int64_t ldv, float scale, void* absmax, void* vt_hi16,
                    void* vt_lo16, int64_t ldvt, int T_pad, void* stream) {
  return split_rhs16(V, m, T, ldv, scale, static_cast<uint32_t*>(absmax), vt_hi16, vt_lo16, ldvt, T_pad,
                     static_cast<cudaStream_t>(stream));
}
int odf_gauss_mmv_pair(int kind, const void* r_hi, const void* r_lo, const float* r_sqnorm, const float* r_opscale,
                       int64_t n_rows, const void* q_hi, const void* q_lo, const float* q_sqnorm, const float* q_opscale,
                       int64_t n_cols, int64_t d, const void* vt_hi16, const void* vt_lo16, int64_t ldvt,
                       const void* v_absmax, int T_pad, int n_splits, float sigma, float* partial, void* panel16,
                       void* stream) {
  if (!(sigma > 0.f)) return set_error(ODF_ERR_ARG, "sigma must be positive");
  if (kind != KIND_TF32 && kind != KIND_F16) return set_error(ODF_ERR_ARG, "unknown operand kind");
  if (!tile2_rows_eligible(n_rows)) return set_error(ODF_ERR_ARG, "too few rows for the CTA-pair tile (odf_tile_pair_eligible)");
  TileLaunch L{};
  L.kind = kind;
  L.r_hi = r_hi; L.r_lo = r_lo; L.r_norm = r_sqnorm; L.r_scale = r_opscale; L.n_rows = n_rows;
  L.q_hi = q_hi; L.q_lo = q_lo; L.q_norm = q_sqnorm; L.q_scale = q_opscale; L.n_cols = n_cols;
  L.d_pad = round_up(d, kblock_elems(kind)); L.T_pad = T_pad;
  L.vt16_hi = vt_hi16; L.vt16_lo = vt_lo16; L.ldvt16 = ldvt; L.v_absmax = static_cast<const uint32_t*>(v_absmax);
  L.mode = MODE_MMV; L.n_splits = n_splits; L.sigma = sigma;
  L.out = partial; L.ldo = T_pad; L.split_stride = n_rows * T_pad;
  L.panel16 = panel16;
  return launch_gauss_tile2(L, static_cast<cudaStream_t>(stream));
}

size_t odf_panel16_bytes(int64_t n_rows, int64_t n_cols) { return panel16_bytes(n_rows, n_cols); }
int odf_panel16_splits(int64_t n_rows, int64_t M) { return panel16_splits(n_rows, M); }
int odf_gauss_mmv_prepared_spill16(int kind, const void* r_hi, const void* r_lo, const float* r_sqnorm,
                                   const float* r_opscale, int64_t n_rows, const void* q_hi, const void* q_lo,
                                   const float* q_sqnorm, const float* q_opscale, int64_t n_cols, int64_t d,
                                   const float* vt_hi, const float* vt_lo, int64_t ldvt, int T_pad, int n_splits,
                                   float sigma, float* partial, void* panel16, void* stream) {
  if (!(sigma > 0.f)) return set_error(ODF_ERR_ARG, "sigma must be positive");
  if (kind != KIND_TF32 && kind != KIND_F16) return set_error(ODF_ERR_ARG, "unknown operand kind");
  if (panel16 == nullptr) return set_error(ODF_ERR_ARG, "panel16 must not be NULL");
  TileLaunch L{};
  L.kind = kind;
  L.r_hi = r_hi; L.r_lo = r_lo; L.r_norm = r_sqnorm; L.r_scale = r_opscale; L.n_rows = n_rows;
  L.q_hi = q_hi; L.q_lo = q_lo; L.q_norm = q_sqnorm; L.q_scale = q_opscale; L.n_cols = n_cols;
  L.d_pad = round_up(d, kblock_elems(kind)); L.vt_hi = vt_hi; L.vt_lo = vt_lo; L.ldvt = ldvt; L.T_pad = T_pad;
  L.mode = MODE_MMV; L.n_splits = n_splits; L.sigma = sigma;
  L.out = partial; L.ldo = T_pad; L.split_stride = n_rows * T_pad;
  L.panel16 = panel16;
  return launch_gauss_tile(L, static_cast<cudaStream_t>(stream));
}
int odf_finish_w16(const float* partial, int n_splits, int64_t n_rows, int T_pad, int64_t T, const float* addend,
                   int64_t ld_add, float* w_f32, void* absmax, void* w16, void* stream) {
  return finish_w16(partial, n_splits, n_rows, T_pad, T, addend, ld_add, w_f32, static_cast<uint32_t*>(absmax), w16,
                    static_cast<cudaStream_t>(stream));
}
int odf_panel16_tmm(const void* panel16, int64_t n_rows, int64_t M, const void* w16, const void* absmax, int T_pad,
                    int n_splits, float* out_partial, void* stream) {
  return launch_panel16_tmm(panel16, n_rows, M, w16, static_cast<const uint32_t*>(absmax), T_pad, n_splits, out_partial,
                            static_cast<cudaStream_t>(stream));
}
int odf_panel16_tmm_hi(const void* panel16, int64_t n_rows, int64_t M, const void* w16, const void* absmax, int T_pad,
                       int n_splits, float* out_partial, void* stream) {
  return launch_panel16_tmm(panel16, n_rows, M, w16, static_cast<const uint32_t*>(absmax), T_pad, n_splits, out_partial,
                            static_cast<cudaStream_t>(stream), 1);
}
int odf_panel16_mmv_hi(const void* panel16, int64_t n_rows, int64_t M, const void* v16, const void* absmax, int T_pad,
                       int n_splits, float* out_partial, void* stream) {
  return launch_panel16_mmv(panel16, n_rows, M, v16, static_cast<const uint32_t*>(absmax), T_pad, n_splits, out_partial,
                            static_cast<cudaStream_t>(stream), 1);
}
int odf_panel16_mmv_splits(int64_t n_rows, int64_t M) { return panel16_mmv_splits(n_rows, M); }
int odf_panel16_mmv(const void* panel16, int64_t n_rows, int64_t M, const void* v16, const void* absmax, int T_pad,
                    int n_splits, float* out_partial, void* stream) {
  return launch_panel16_mmv(panel16, n_rows, M, v16, static_cast<const uint32_t*>(absmax), T_pad, n_splits, out_partial,
                            static_cast<cudaStream_t>(stream));
}

int odf_finish_rows(const float* partial, int n_splits, int64_t n_rows, int T_pad, int64_t T, float scale,
                    const float* addend, int64_t ld_add, float* out, int64_t ldo, void* stream) {
  return finish_rows(partial, n_splits, n_rows, T_pad, T, scale, addend, ld_add, out, ldo, static_cast<cudaStream_t>(stream));
}
int odf_finish_split(const float* partial, int n_splits, int64_t n_rows, int T_pad, int64_t T, float scale,
                     const float* addend, int64_t ld_add, float* wt_hi, float* wt_lo, int64_t ldwt, void* stream) {
  return finish_split(partial, n_splits, n_rows, T_pad, T, scale, addend, ld_add, wt_hi, wt_lo, ldwt, static_cast<cudaStream_t>(stream));
}

int odf_gauss_kmm_prepared(int kind, const void* c_hi, const void* c_lo, const float* c_sqnorm,
                           const float* c_opscale, int64_t M, int64_t d, float sigma, float* K, int64_t ldk,
                           void* stream) {
  if (!(sigma > 0.f)) return set_error(ODF_ERR_ARG, "sigma must be positive");
  if (kind != KIND_TF32 && kind != KIND_F16) return set_error(ODF_ERR_ARG, "unknown operand kind");
  TileLaunch L{};
  L.kind = kind;
  L.r_hi = c_hi; L.r_lo = c_lo; L.r_norm = c_sqnorm; L.r_scale = c_opscale; L.n_rows = M;
  L.q_hi = c_hi; L.q_lo = c_lo; L.q_norm = c_sqnorm; L.q_scale = c_opscale; L.n_cols = M;
  L.d_pad = round_up(d, kblock_elems(kind)); L.T_pad = 16; L.mode = MODE_STORE;
  L.n_splits = odf_tile_splits(M, M, d, kind); L.sigma = sigma;
  L.out = K; L.ldo = ldk; L.split_stride = 0;
  return launch_gauss_tile(L, static_cast<cudaStream_t>(stream));
}

int odf_gemm_nt_split(int kind, const void* a_hi, const void* a_lo, const float* a_sqnorm, const float* a_opscale,
                      int64_t m, const void* b_hi, const void* b_lo, const float* b_sqnorm, const float* b_opscale,
                      int64_t n, int64_t k, float alpha, float beta, float* C, int64_t ldc, void* stream) {
  if (kind != KIND_TF32 && kind != KIND_F16) return set_error(ODF_ERR_ARG, "unknown operand kind");
  if (m <= 0 || n <= 0 || k <= 0 || ldc < n) return set_error(ODF_ERR_ARG, "gemm_nt_split: bad shape");
  TileLaunch L{};
  L.kind = kind;
  L.r_hi = a_hi; L.r_lo = a_lo; L.r_norm = a_sqnorm; L.r_scale = a_opscale; L.n_rows = m;
  L.q_hi = b_hi; L.q_lo = b_lo; L.q_norm = b_sqnorm; L.q_scale = b_opscale; L.n_cols = n;
  L.d_pad = round_up(k, kblock_elems(kind)); L.T_pad = 16; L.mode = MODE_STORE;
  L.n_splits = odf_tile_splits(m, n, k, kind); L.sigma = 1.f;
  L.out = C; L.ldo = ldc; L.split_stride = 0;
  L.linear = 1; L.lin_alpha = alpha; L.lin_beta = beta;
  return launch_gauss_tile(L, static_cast<cudaStream_t>(stream));
}

// Scratch odf_precond_init / odf_potrf_upper need for an M x M factorisation, as cuSOLVER reports it for the current
// device (0 when no device / handle can be had: callers then fall back to the closed form in odf_workspace_bytes).
size_t odf_precond_workspace_bytes(int64_t M) {
  if (M <= 0 || M > 0x7fffffff) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return 0;
  }
  if (ensure_handles(nullptr) != ODF_OK) return 0;
  int lwork = 0;
  if (cusolverDnSpotrf_bufferSize(g_cusolver, CUBLAS_FILL_MODE_LOWER, static_cast<int>(M), nullptr, static_cast<int>(M), &lwork) !=
      CUSOLVER_STATUS_SUCCESS)
    return 0;
  return al(sizeof(int) * 2) + al(sizeof(float) * static_cast<size_t>(lwork)) + 256;
}

// ---------------------------------------------------------------- convenience (plain fp32 in)
size_t odf_workspace_bytes(int op, int64_t n, int64_t M, int64_t d, int64_t T) {
  const int T_pad = tpad_of(T > 0 ? T : 1);
  const int kind = g_default_kind;
  switch (op) {
    case ODF_OP_MMV: {
      const int S = odf_tile_splits(n, M, d, kind);
      return prepared_bytes(n, d, kind) + prepared_bytes(M, d, kind) + 2 * al(sizeof(float) * T_pad * round_up(M, 128)) +
             al(sizeof(float) * S * n * T_pad) + al(128);
    }
    case ODF_OP_DMMV: {
      const int S1 = odf_tile_splits(n, M, d, kind);
      const int S2 = odf_tile_splits(M, n, d, kind);
      return prepared_bytes(n, d, kind) + prepared_bytes(M, d, kind) + 2 * al(sizeof(float) * T_pad * round_up(M, 128)) +
             2 * al(sizeof(float) * T_pad * round_up(n, 128)) + al(sizeof(float) * S1 * n * T_pad) +
             al(sizeof(float) * S2 * M * T_pad) + al(sizeof(float) * n * T_pad) + 2 * al(128);
    }
    case ODF_OP_KMM:
      return prepared_bytes(M, d, kind);
    case ODF_OP_PRECOND: {
      // potrf scratch: asked from cuSOLVER when a device is current (odf_precond_workspace_bytes); this closed form
      // is the lower bound used when it cannot be queried
      const size_t q = odf_precond_workspace_bytes(M);
      const size_t lb = al(sizeof(float) * (static_cast<size_t>(M) * 256 + (1u << 20))) + 256;
      return q > lb ? q : lb;
    }
    default:
      return 0;
  }
}

// partial[s] = K(rows, cols restricted to split s) . V for a plain fp32 V (m x T): splits V into the operand format of
// the tile that will run (fp16 pairs + per-column scales for the CTA-pair tile when the launch has >= 8192 rows, tf32
// hi / lo otherwise) inside the two scratch arrays vbuf_a / vbuf_b (each T_pad x ldvt floats) and launches it.
static int mmv_from_plain(int kind, const Prepared& rows, int64_t n_rows, const Prepared& cols, int64_t n_cols, int64_t d,
                          const float* V, int64_t T, int64_t ldv, int T_pad, int64_t ldvt, float* vbuf_a, float* vbuf_b,
                          uint32_t* absmax, int S, float sigma, float* partial, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if (tile2_rows_eligible(n_rows)) {
    if ((rc = split_rhs16(V, n_cols, T, ldv, 1.f, absmax, vbuf_a, vbuf_b, ldvt, T_pad, st))) return rc;
    return odf_gauss_mmv_pair(kind, rows.hi, rows.lo, rows.sqn, rows.opscale, n_rows, cols.hi, cols.lo, cols.sqn,
                              cols.opscale, n_cols, d, vbuf_a, vbuf_b, ldvt, absmax, T_pad, S, sigma, partial, nullptr, stream);
  }
  if ((rc = split_rhs(V, n_cols, T, ldv, 1.f, vbuf_a, vbuf_b, ldvt, T_pad, st))) return rc;
  return odf_gauss_mmv_prepared(kind, rows.hi, rows.lo, rows.sqn, rows.opscale, n_rows, cols.hi, cols.lo, cols.sqn,
                                cols.opscale, n_cols, d, vbuf_a, vbuf_b, ldvt, T_pad, S, sigma, partial, stream);
}

int odf_gauss_mmv(const float* X, int64_t n, int64_t ldx, const float* C, int64_t M, int64_t ldc, int64_t d,
                  const float* V, int64_t T, int64_t ldv, float sigma, float* out, int64_t ldo, void* ws,
                  size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T_pad = tpad_of(T);
  if (T_pad < 0) return set_error(ODF_ERR_ARG, "T must be <= 32 per call");
  const int64_t ldvt = round_up(M, 128);
  const int kind = g_default_kind;
  Arena a(ws, ws_bytes);
  Prepared px, pc;
  take_prepared(a, n, d, kind, &px);
  take_prepared(a, M, d, kind, &pc);
  float* vth = a.take<float>(T_pad * ldvt);
  float* vtl = a.take<float>(T_pad * ldvt);
  const int S = odf_tile_splits(n, M, d, kind);
  float* partial = a.take<float>(static_cast<size_t>(S) * n * T_pad);
  uint32_t* absmax = a.take<uint32_t>(32);
  if (!a.ok()) return set_error(ODF_ERR_WORKSPACE, "odf_gauss_mmv: workspace too small");
  int rc;
  if ((rc = prepare_points(X, n, d, ldx, nullptr, 1.f, kind, px.hi, px.lo, px.sqn, px.opscale, st))) return rc;
  if ((rc = prepare_points(C, M, d, ldc, nullptr, 1.f, kind, pc.hi, pc.lo, pc.sqn, pc.opscale, st))) return rc;
  if ((rc = mmv_from_plain(kind, px, n, pc, M, d, V, T, ldv, T_pad, ldvt, vth, vtl, absmax, S, sigma, partial, stream))) return rc;
  return finish_rows(partial, S, n, T_pad, T, 1.f, nullptr, 0, out, ldo, st);
}

int odf_gauss_dmmv(const float* X, int64_t n, int64_t ldx, const float* C, int64_t M, int64_t ldc, int64_t d,
                   const float* V, int64_t ldv, const float* W, int64_t ldw, int64_t T, float sigma, float* out,
                   int64_t ldo, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T_pad = tpad_of(T);
  if (T_pad < 0) return set_error(ODF_ERR_ARG, "T must be <= 32 per call");
  if (!V && !W) return set_error(ODF_ERR_ARG, "dmmv needs V or W");
  const int64_t ldvt = round_up(M, 128), ldwt = round_up(n, 128);
  const int kind = g_default_kind;
  Arena a(ws, ws_bytes);
  Prepared px, pc;
  take_prepared(a, n, d, kind, &px);
  take_prepared(a, M, d, kind, &pc);
  float* vth = a.take<float>(T_pad * ldvt);
  float* vtl = a.take<float>(T_pad * ldvt);
  float* wth = a.take<float>(T_pad * ldwt);
  float* wtl = a.take<float>(T_pad * ldwt);
  const int S1 = odf_tile_splits(n, M, d, kind);
  const int S2 = odf_tile_splits(M, n, d, kind);
  float* part1 = a.take<float>(static_cast<size_t>(S1) * n * T_pad);
  float* part2 = a.take<float>(static_cast<size_t>(S2) * M * T_pad);
  float* Wf = a.take<float>(static_cast<size_t>(n) * T_pad);
  uint32_t* absmax_v = a.take<uint32_t>(32);
  uint32_t* absmax_w = a.take<uint32_t>(32);
  if (!a.ok()) return set_error(ODF_ERR_WORKSPACE, "odf_gauss_dmmv: workspace too small");
  int rc;
  if ((rc = prepare_points(X, n, d, ldx, nullptr, 1.f, kind, px.hi, px.lo, px.sqn, px.opscale, st))) return rc;
  if ((rc = prepare_points(C, M, d, ldc, nullptr, 1.f, kind, pc.hi, pc.lo, pc.sqn, pc.opscale, st))) return rc;
  const float* Wsrc = W;
  int64_t ldwsrc = ldw;
  if (V) {
    // first half: W' = K V + W  (n x T)
    if ((rc = mmv_from_plain(kind, px, n, pc, M, d, V, T, ldv, T_pad, ldvt, vth, vtl, absmax_v, S1, sigma, part1, stream))) return rc;
    if ((rc = finish_rows(part1, S1, n, T_pad, T, 1.f, W, ldw, Wf, T_pad, st))) return rc;
    Wsrc = Wf;
    ldwsrc = T_pad;
  }
  // second half: K^T W'  (rows = centres, columns = data)
  if ((rc = mmv_from_plain(kind, pc, M, px, n, d, Wsrc, T, ldwsrc, T_pad, ldwt, wth, wtl, absmax_w, S2, sigma, part2, stream))) return rc;
  return finish_rows(part2, S2, M, T_pad, T, 1.f, nullptr, 0, out, ldo, st);
}

int odf_gauss_kmm(const float* C, int64_t M, int64_t ldc, int64_t d, float sigma, float* K, int64_t ldk, void* ws,
                  size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int kind = g_default_kind;
  Arena a(ws, ws_bytes);
  Prepared pc;
  if (!take_prepared(a, M, d, kind, &pc)) return set_error(ODF_ERR_WORKSPACE, "odf_gauss_kmm: workspace too small");
  int rc;
  if ((rc = prepare_points(C, M, d, ldc, nullptr, 1.f, kind, pc.hi, pc.lo, pc.sqn, pc.opscale, st))) return rc;
  return odf_gauss_kmm_prepared(kind, pc.hi, pc.lo, pc.sqn, pc.opscale, M, d, sigma, K, ldk, stream);
}

// ---------------------------------------------------------------- preconditioner
// Row-major upper U  ==  column-major lower L = U^T, so cuSOLVER/cuBLAS run with uplo = LOWER.
int odf_precond_init(float* Tm, float* Am, int64_t M, float lam, float eps, void* ws, size_t ws_bytes,
                     void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (M <= 0 || M > 0x7fffffff) return set_error(ODF_ERR_ARG, "precond_init: bad M");
  int rc;
  if ((rc = ensure_handles(st))) return rc;
  const int m = static_cast<int>(M);
  int lwork = 0;
  if (cusolverDnSpotrf_bufferSize(g_cusolver, CUBLAS_FILL_MODE_LOWER, m, Tm, m, &lwork) != CUSOLVER_STATUS_SUCCESS)
    return set_error(ODF_ERR_CUDA, "cusolverDnSpotrf_bufferSize failed");
  Arena a(ws, ws_bytes);
  int* info = a.take<int>(2);
  float* work = a.take<float>(static_cast<size_t>(lwork));
  if (!a.ok()) return set_error(ODF_ERR_WORKSPACE, "precond_init: workspace too small");

  // T: K_MM + eps*M*I = L L^T
  if ((rc = add_diag(Tm, M, eps * static_cast<float>(M), st))) return rc;
  if (cusolverDnSpotrf(g_cusolver, CUBLAS_FILL_MODE_LOWER, m, Tm, m, work, lwork, info) != CUSOLVER_STATUS_SUCCESS)
    return set_error(ODF_ERR_CUDA, "cusolverDnSpotrf (T) failed to launch");
  if ((rc = zero_lower(Tm, M, st))) return rc;  // row-major strict lower == the triangle potrf left untouched
  // A: (1/M) T T^T + lam I = (1/M) L^T L + lam I
  if ((rc = ttt_upper(Tm, Am, M, 1.f / static_cast<float>(M)))) return rc;
  if ((rc = add_diag(Am, M, lam, st))) return rc;
  if (cusolverDnSpotrf(g_cusolver, CUBLAS_FILL_MODE_LOWER, m, Am, m, work, lwork, info + 1) != CUSOLVER_STATUS_SUCCESS)
    return set_error(ODF_ERR_CUDA, "cusolverDnSpotrf (A) failed to launch");
  if ((rc = zero_lower(Am, M, st))) return rc;
  int hinfo[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(hinfo, info, sizeof hinfo, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_cuda_error(e, "precond_init");
  if (hinfo[0] != 0 || hinfo[1] != 0) {
    char buf[160];
    snprintf(buf, sizeof buf, "Cholesky failed: info(T)=%d info(A)=%d (matrix not positive definite)", hinfo[0], hinfo[1]);
    return set_error(ODF_ERR_LINALG, buf);
  }
  return ODF_OK;
}

// Tensor-core build (odf_precond.cu): K_MM in, T / A and (optionally) their explicit inverses out, all upper triangular.
size_t odf_precond_build_workspace_bytes(int64_t M) {
  if (M <= 0 || M > 0x7fffffff) return 0;
  int lwork = 0, ndev = 0;
  if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && ensure_handles(nullptr) == ODF_OK) {
    const int nb = static_cast<int>(M < 4096 ? M : 4096);
    if (cusolverDnSpotrf_bufferSize(g_cusolver, CUBLAS_FILL_MODE_UPPER, nb, nullptr, static_cast<int>(M), &lwork) != CUSOLVER_STATUS_SUCCESS)
      lwork = 0;
  } else {
    cudaGetLastError();
  }
  if (lwork < (1 << 22)) lwork = 1 << 22;                  // generous floor: the exact figure is re-checked inside the build
  return precond_build_workspace_bytes(M, lwork);
}
int odf_precond_build(float* K, float* Tm, float* Am, float* Tinv, float* Ainv, int64_t M, float lam, float eps, void* ws,
                      size_t ws_bytes, void* stream) {
  return precond_build(K, Tm, Am, Tinv, Ainv, M, lam, eps, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

/* Building blocks of the same preconditioner for the row-sharded multi-GPU fit, where the O(M^3) pieces are
 * split over the ranks (odf/falkon.py::_build_preconditioner): in-place Cholesky of a row-major symmetric
 * matrix into its upper factor, diagonal shift, triangle clear. */
int odf_potrf_upper(float* A, int64_t M, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (M <= 0 || M > 0x7fffffff) return set_error(ODF_ERR_ARG, "potrf_upper: bad M");
  int rc;
  if ((rc = ensure_handles(st))) return rc;
  const int m = static_cast<int>(M);
  int lwork = 0;
  if (cusolverDnSpotrf_bufferSize(g_cusolver, CUBLAS_FILL_MODE_LOWER, m, A, m, &lwork) != CUSOLVER_STATUS_SUCCESS)
    return set_error(ODF_ERR_CUDA, "cusolverDnSpotrf_bufferSize failed");
  Arena a(ws, ws_bytes);
  int* info = a.take<int>(2);
  float* work = a.take<float>(static_cast<size_t>(lwork));
  if (!a.ok()) return set_error(ODF_ERR_WORKSPACE, "potrf_upper: workspace too small");
  // row-major upper U == column-major lower L = U^T
  if (cusolverDnSpotrf(g_cusolver, CUBLAS_FILL_MODE_LOWER, m, A, m, work, lwork, info) != CUSOLVER_STATUS_SUCCESS)
    return set_error(ODF_ERR_CUDA, "cusolverDnSpotrf failed to launch");
  if ((rc = zero_lower(A, M, st))) return rc;
  int hinfo = 0;
  cudaError_t e = cudaMemcpyAsync(&hinfo, info, sizeof hinfo, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_cuda_error(e, "potrf_upper");
  if (hinfo != 0) {
    char buf[128];
    snprintf(buf, sizeof buf, "Cholesky failed: info=%d (matrix not positive definite)", hinfo);
    return set_error(ODF_ERR_LINALG, buf);
  }
  return ODF_OK;
}
int odf_add_diag(float* A, int64_t M, float value, void* stream) { return add_diag(A, M, value, static_cast<cudaStream_t>(stream)); }
int odf_zero_strict_lower(float* A, int64_t M, void* stream) { return zero_lower(A, M, static_cast<cudaStream_t>(stream)); }

int odf_precond_solve(const float* Tri, int64_t M, float* B, int64_t T, int64_t ldb, int which, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if ((rc = ensure_handles(st))) return rc;
  if (which < 0 || which > 3) return set_error(ODF_ERR_ARG, "precond_solve: bad `which`");
  // Row-major B (M x T) is column-major B' = B^T (T x M).  X = U^-1 B  <=>  X' L = B' (L = U^T col-major lower);
  // X = U^-T B  <=>  X' L^T = B'.
  const bool transposed = (which == ODF_SOLVE_TT || which == ODF_SOLVE_AT);
  const float one = 1.f;
  cublasStatus_t s = cublasStrsm(g_cublas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER,
                                 transposed ? CUBLAS_OP_T : CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, static_cast<int>(T),
                                 static_cast<int>(M), &one, Tri, static_cast<int>(M), B, static_cast<int>(ldb));
  if (s != CUBLAS_STATUS_SUCCESS) return set_error(ODF_ERR_CUDA, "cublasStrsm failed");
  return ODF_OK;
}

int odf_precond_invert(const float* Tri, float* Inv, int64_t M, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if ((rc = ensure_handles(st))) return rc;
  if ((rc = set_identity(Inv, M, st))) return rc;
  if ((rc = inv_upper_rec(Tri, Inv, M, M, st))) return rc;
  return zero_lower(Inv, M, st);      // exact zeros below the diagonal (scratch blocks and TRSM rounding dust)
}

// Plain row-major fp32 GEMM C (m x n) = alpha op(A) op(B) + beta C on sub-blocks (pitches lda / ldb / ldc): the
// building block of the distributed preconditioner (triangle-aware column blocks of T T^T).
int odf_gemm(int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda,
             const float* B, int64_t ldb, float beta, float* C, int64_t ldc, void* stream) {
  int rc;
  if ((rc = ensure_handles(static_cast<cudaStream_t>(stream)))) return rc;
  if (m <= 0 || n <= 0 || k <= 0) return set_error(ODF_ERR_ARG, "odf_gemm: empty product");
  return gemm_rm(trans_a != 0, trans_b != 0, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

// Rows [r0, r1) of op(Inv) Bin for an UPPER-triangular Inv: only the non-zero part of the operand is read
// (plain: columns >= r0; transposed: rows < r1).  The row-sharded fit gives every rank one row block and all-gathers.
int odf_precond_apply_rows(const float* Inv, int64_t M, int64_t r0, int64_t r1, const float* Bin, float* Bout_rows,
                           int64_t T, int64_t ldb, int64_t ldo, int transposed, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (r0 < 0 || r1 > M || r0 >= r1 || T <= 0 || ldb < T || ldo < T) return set_error(ODF_ERR_ARG, "precond_apply_rows: bad shape");
  int rc;
  for (int64_t t0 = 0; t0 < T; t0 += 32) {
    const int64_t tw = (T - t0 < 32) ? (T - t0) : 32;
    if ((rc = tri_apply(Inv, M, Bin + t0, ldb, tw, Bout_rows + t0, ldo, r0, r1, transposed, st))) return rc;
  }
  return ODF_OK;
}

// Bout = op(Inv) Bin for an upper-triangular Inv: own kernel (csrc/odf_tri.cu) that reads only the triangle and accumulates
// in fp64; ODF_PRECOND_APPLY=cublas selects round 1's full-square cuBLAS sgemm.
int odf_precond_apply(const float* Inv, int64_t M, const float* Bin, float* Bout, int64_t T, int64_t ldb,
                      int transposed, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Bin == Bout) return set_error(ODF_ERR_ARG, "precond_apply is out of place");
  static int use_cublas = -1;
  if (use_cublas < 0) {
    const char* e = getenv("ODF_PRECOND_APPLY");
    use_cublas = (e && strcmp(e, "cublas") == 0) ? 1 : 0;
  }
  if (!use_cublas) return odf_precond_apply_rows(Inv, M, 0, M, Bin, Bout, T, ldb, ldb, transposed, stream);
  int rc;
  if ((rc = ensure_handles(st))) return rc;
  // row-major Bout = op(Inv) Bin   <=>   column-major Bout' (T x M) = Bin' (T x M) * op'(Inv')
  const float one = 1.f, zero = 0.f;
  cublasStatus_t s = cublasSgemm(g_cublas, CUBLAS_OP_N, transposed ? CUBLAS_OP_T : CUBLAS_OP_N, static_cast<int>(T),
                                 static_cast<int>(M), static_cast<int>(M), &one, Bin, static_cast<int>(ldb), Inv,
                                 static_cast<int>(M), &zero, Bout, static_cast<int>(ldb));
  if (s != CUBLAS_STATUS_SUCCESS) return set_error(ODF_ERR_CUDA, "cublasSgemm (precond_apply) failed");
  return ODF_OK;
}

// ---------------------------------------------------------------- RLS box refiners (odf_rls.cu)
size_t odf_rls_workspace_bytes(int64_t n, int64_t d, int64_t n_classes) {
  if (n <= 0 || d <= 0 || n_classes <= 0) return 0;
  int lwork = 0, ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    lwork = 1 << 22;
  } else if (rls_query_lwork(d, &lwork) != ODF_OK) {
    lwork = 1 << 22;
  }
  return rls_workspace_bytes(n, d, n_classes, lwork);
}
int odf_rls_train(const float* X, int64_t n, int64_t d, int64_t ldx, const double* Yw, const int64_t* perm, const int64_t* seg_host,
                  const int* row_class, int64_t n_classes, double lam, float* W, float* losses, void* ws, size_t ws_bytes,
                  void* stream) {
  return rls_train(X, n, d, ldx, Yw, perm, seg_host, row_class, n_classes, lam, W, losses, ws, ws_bytes,
                   static_cast<cudaStream_t>(stream));
}
int odf_rls_apply(const float* feat, int64_t n, int64_t d, int64_t ldf, const float* Wp, const float* bias, const float* Tinv,
                  const float* mu, const float* ex_boxes, int64_t C, float img_w, float img_h, float eps, const float* mean,
                  float zscale, float* out, void* stream) {
  return rls_apply(feat, n, d, ldf, Wp, bias, Tinv, mu, ex_boxes, C, img_w, img_h, eps, mean, zscale, out,
                   static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------- CG vector kernels
size_t odf_cg_workspace_bytes(int64_t M, int64_t T) { return cg_workspace_bytes(M, T); }
int odf_cg_init(const float* R, int64_t M, int64_t T, int64_t ld, float* state, void* ws, size_t ws_bytes, void* stream) {
  return cg_init(R, M, T, ld, state, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int odf_cg_alpha(const float* P, const float* AP, int64_t M, int64_t T, int64_t ld, float eps, float* state, void* ws,
                 size_t ws_bytes, void* stream) {
  return cg_alpha(P, AP, M, T, ld, eps, state, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int odf_cg_axpy_a(float* Y, const float* X, int64_t M, int64_t T, int64_t ld, float sign, const float* state, void* stream) {
  return cg_axpy_a(Y, X, M, T, ld, sign, state, static_cast<cudaStream_t>(stream));
}
int odf_cg_residual(float* R, const float* Bm, const float* H, int64_t M, int64_t T, int64_t ld, const float* state,
                    void* stream) {
  return cg_residual(R, Bm, H, M, T, ld, state, static_cast<cudaStream_t>(stream));
}
int odf_cg_beta(const float* R, int64_t M, int64_t T, int64_t ld, float eps, float tol, float* state, void* ws,
                size_t ws_bytes, void* stream) {
  return cg_beta(R, M, T, ld, eps, tol, state, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int odf_cg_xpby_b(float* P, const float* R, int64_t M, int64_t T, int64_t ld, const float* state, void* stream) {
  return cg_xpby_b(P, R, M, T, ld, state, static_cast<cudaStream_t>(stream));
}
int odf_axpby(float* out, float alpha, const float* A, float beta, const float* Bm, int64_t M, int64_t T, int64_t ld,
              void* stream) {
  return axpby(out, alpha, A, beta, Bm, M, T, ld, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
