// RLS box-refinement regressors, all classes / anchors in one batched path (fp64, as the reference computes them):
//   train  src/modules/region-refiner/region_refiner_trainer/train_region_refiner.py:25-119
//          per class c over its rows:  Xi = [X 1],  w_k = (Xi^T Xi + lam I)^-1 Xi^T y'_k  (k < 4, y' = whitened targets),
//          losses = (Xi w_k - y'_k)^2 / 2
//   apply  src/modules/region-refiner/region_predictor/predict_regions.py:16-80  (also the *_parallel heads,
//          roi_box_predictors.py:97-124, rpn.py:158-187):  Y = feat W + b -> un-whiten (Y T_inv + mu) -> box decode
//
// Training cost is the normal matrix: n_c (d+1)^2 fp64 flops per class (d + 1 up to 2049).  rls_gram_kernel forms
//     G_c = Z_c^T Z_c,   Z = [X | 1 | y'_0..y'_3]            (upper 64 x 64 tiles only)
// for every class in ONE launch on the fp64 tensor cores (mma.sync m8n8k4 f64: tcgen05 has no fp64 kind), reading the
// fp32 features through the class-sorted row permutation (no gathered copy of X) and widening them on the fly; the last
// four columns of G_c are the right-hand sides Xi^T y'.  The (d+1)^3/3 factorisations go to cuSOLVER Dpotrf / Dpotrs,
// round-robin over side streams (independent, latency-bound).  The losses are one more pass over the rows.
#include <cstdlib>
#include <vector>

#include <cusolverDn.h>

#include "odf_internal.h"

namespace odf {

namespace {

constexpr int GT = 64;          // Gram tile
constexpr int GK = 16;          // rows of Z per shared-memory chunk
constexpr int GP = GT + 4;      // row pitch (doubles): 8 (mod 32) words -> conflict-free fragment loads
constexpr int RLS_STREAMS = 4;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// element (row r of the class segment, column c of Z = [X | 1 | y']) as a double; zero outside
__device__ __forceinline__ double z_elem(const float* __restrict__ X, int64_t ldx, int d, const double* __restrict__ Yw,
                                         int64_t grow, int c) {
  if (c < d) return static_cast<double>(__ldg(X + grow * ldx + c));
  if (c == d) return 1.0;
  if (c < d + 5) return Yw[grow * 4 + (c - d - 1)];
  return 0.0;
}

// grid: (upper tile pairs, 1, classes); 128 threads = 4 warps, each a 32 x 32 quarter of the 64 x 64 tile
__global__ void __launch_bounds__(128)
rls_gram_kernel(const float* __restrict__ X, int64_t ldx, int d, const double* __restrict__ Yw,
                const int64_t* __restrict__ perm, const int64_t* __restrict__ seg, int n_tiles, int64_t P,
                double* __restrict__ G) {
  __shared__ double As[GK][GP];
  __shared__ double Bs[GK][GP];
  __shared__ int64_t rows[GK];
  const int cls = blockIdx.z;
  // linear index -> (ti <= tj)
  int tj = static_cast<int>((sqrtf(8.f * static_cast<float>(blockIdx.x) + 1.f) - 1.f) * 0.5f);
  while ((tj + 1) * (tj + 2) / 2 <= static_cast<int>(blockIdx.x)) ++tj;
  while (tj * (tj + 1) / 2 > static_cast<int>(blockIdx.x)) --tj;
  const int ti = static_cast<int>(blockIdx.x) - tj * (tj + 1) / 2;
  (void)n_tiles;
  const int64_t s0 = seg[cls], s1 = seg[cls + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (warp >> 1) * 32, n0 = (warp & 1) * 32;
  const int lk = lane & 3, lm = lane >> 2;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  const int lr = threadIdx.x >> 3, lc = (threadIdx.x & 7) * 8;       // loader: row lr, 8 columns from lc
  const bool vec = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  for (int64_t r0 = s0; r0 < s1; r0 += GK) {
    if (threadIdx.x < GK) rows[threadIdx.x] = (r0 + threadIdx.x < s1) ? perm[r0 + threadIdx.x] : -1;
    __syncthreads();
    const int64_t grow = rows[lr];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int c0 = (half == 0 ? ti : tj) * GT + lc;
      double* dst = (half == 0 ? &As[lr][lc] : &Bs[lr][lc]);
      if (grow < 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[k] = 0.0;
      } else if (vec && c0 + 8 <= d) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(X + grow * ldx + c0));
        const float4 v = __ldg(reinterpret_cast<const float4*>(X + grow * ldx + c0 + 4));
        dst[0] = u.x; dst[1] = u.y; dst[2] = u.z; dst[3] = u.w;
        dst[4] = v.x; dst[5] = v.y; dst[6] = v.z; dst[7] = v.w;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[k] = z_elem(X, ldx, d, Yw, grow, c0 + k);
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        a[t] = As[kk + lk][m0 + t * 8 + lm];
        b[t] = Bs[kk + lk][n0 + t * 8 + lm];
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
    __syncthreads();
  }
  double* Gc = G + static_cast<int64_t>(cls) * P * P;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int64_t r = static_cast<int64_t>(ti) * GT + m0 + mt * 8 + lm;
      const int64_t c = static_cast<int64_t>(tj) * GT + n0 + nt * 8 + lk * 2;
      *reinterpret_cast<double2*>(Gc + r * P + c) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
}

// G_c[i][i] += lam for i <= d;  B_c (column-major [(d+1) x 4], i.e. row-major [4][d+1]) <- G_c[0:d+1, d+1:d+5]
__global__ void __launch_bounds__(256)
rls_rhs_kernel(double* __restrict__ G, int64_t P, int d, double lam, double* __restrict__ Bm) {
  const int cls = blockIdx.y;
  double* Gc = G + static_cast<int64_t>(cls) * P * P;
  double* Bc = Bm + static_cast<int64_t>(cls) * 4 * (d + 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= d; i += gridDim.x * blockDim.x) {
    Gc[static_cast<int64_t>(i) * P + i] += lam;
#pragma unroll
    for (int k = 0; k < 4; ++k) Bc[static_cast<int64_t>(k) * (d + 1) + i] = Gc[static_cast<int64_t>(i) * P + d + 1 + k];
  }
}

// weights (fp32 out, [classes][4][d+1]) and losses ([n][4] in permuted row order, fp32): one warp per row
__global__ void __launch_bounds__(256)
rls_loss_kernel(const float* __restrict__ X, int64_t ldx, int d, const double* __restrict__ Yw, const int64_t* __restrict__ perm,
                const int* __restrict__ row_class, int64_t n, const double* __restrict__ Bm, float* __restrict__ losses) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const int64_t grow = perm[r];
  const double* w = Bm + static_cast<int64_t>(row_class[r]) * 4 * (d + 1);
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int c = lane; c < d; c += 32) {
    const double x = static_cast<double>(__ldg(X + grow * ldx + c));
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = fma(x, w[static_cast<int64_t>(k) * (d + 1) + c], s[k]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  }
  if (lane < 4) {
    const double res = s[lane] + w[static_cast<int64_t>(lane) * (d + 1) + d] - Yw[grow * 4 + lane];
    losses[r * 4 + lane] = static_cast<float>(0.5 * res * res);
  }
}

__global__ void __launch_bounds__(256)
f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = static_cast<float>(in[i]);
}

// Fused apply: one CTA per RoI.  y[j] = feat . Wp[:, j] + b[j] (j < 4 C, fp32), un-whitening per class (y T_inv + mu), box
// decode with the predictor's eps-width convention, clamp; slot 0 of the output row is the un-refined box.
__global__ void __launch_bounds__(128)
rls_apply_kernel(const float* __restrict__ feat, int64_t ldf, int d, const float* __restrict__ Wp /* [d][4C] */,
                 const float* __restrict__ bias /* [4C] */, const float* __restrict__ Tinv /* [C][4][4] */,
                 const float* __restrict__ mu /* [C][4] */, const float* __restrict__ ex /* [n][4] */, int C, float img_w,
                 float img_h, float eps, const float* __restrict__ mean, float zscale, float* __restrict__ out /* [n][C+1][4] */) {
  extern __shared__ float sm[];
  float* f = sm;                // d
  float* y = sm + d;            // 4C
  const int64_t r = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float v = __ldg(feat + r * ldf + c);
    if (mean) v = (v - __ldg(mean + c)) * zscale;         // zScores fused into the load
    f[c] = v;
  }
  __syncthreads();
  const int J = 4 * C;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < d; ++c) acc = fmaf(f[c], __ldg(Wp + static_cast<int64_t>(c) * J + j), acc);
    y[j] = acc + __ldg(bias + j);
  }
  __syncthreads();
  const float x1 = ex[r * 4 + 0], y1 = ex[r * 4 + 1], x2 = ex[r * 4 + 2], y2 = ex[r * 4 + 3];
  float* orow = out + r * static_cast<int64_t>(C + 1) * 4;
  if (threadIdx.x == 0) { orow[0] = x1; orow[1] = y1; orow[2] = x2; orow[3] = y2; }
  const float sw = x2 - x1 + eps, sh = y2 - y1 + eps;
  const float cx = x1 + 0.5f * sw, cy = y1 + 0.5f * sh;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = __ldg(mu + c * 4 + j);
#pragma unroll
      for (int k = 0; k < 4; ++k) a = fmaf(y[c * 4 + k], __ldg(Tinv + (c * 4 + k) * 4 + j), a);
      u[j] = a;
    }
    const float pcx = u[0] * sw + cx, pcy = u[1] * sh + cy;
    const float pw = expf(u[2]) * sw, ph = expf(u[3]) * sh;
    float* o = orow + (c + 1) * 4;
    o[0] = fmaxf(pcx - 0.5f * pw, 0.f);
    o[1] = fmaxf(pcy - 0.5f * ph, 0.f);
    o[2] = fminf(pcx + 0.5f * pw - 1.f, img_w - 1.f);
    o[3] = fminf(pcy + 0.5f * ph - 1.f, img_h - 1.f);
  }
}

struct RlsLib {
  bool made = false;
  cudaStream_t st[RLS_STREAMS];
  cusolverDnHandle_t sol[RLS_STREAMS];
  cudaEvent_t fork, join[RLS_STREAMS];
};
RlsLib g_rls[kMaxDevices];

int rls_lib(RlsLib** out) {
  RlsLib& R = g_rls[current_device()];
  if (!R.made) {
    bool ok = cudaEventCreateWithFlags(&R.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int s = 0; ok && s < RLS_STREAMS; ++s) {
      ok = cudaStreamCreateWithFlags(&R.st[s], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&R.join[s], cudaEventDisableTiming) == cudaSuccess &&
           cusolverDnCreate(&R.sol[s]) == CUSOLVER_STATUS_SUCCESS;
      if (ok) cusolverDnSetStream(R.sol[s], R.st[s]);
    }
    if (!ok) return set_error(ODF_ERR_CUDA, "rls: cannot create streams / cuSOLVER handles");
    R.made = true;
  }
  *out = &R;
  return ODF_OK;
}

inline size_t al256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }
inline int64_t gram_pitch(int64_t d) { return round_up(d + 5, GT); }

}  // namespace

size_t rls_workspace_bytes(int64_t n, int64_t d, int64_t n_classes, int lwork) {
  const int64_t P = gram_pitch(d);
  (void)n;
  return al256(static_cast<size_t>(n_classes) * P * P * 8) + al256(static_cast<size_t>(n_classes) * 4 * (d + 1) * 8) +
         RLS_STREAMS * al256(static_cast<size_t>(lwork > 0 ? lwork : 1) * 8) + al256(sizeof(int) * static_cast<size_t>(2 * n_classes + 2)) +
         al256(sizeof(int64_t) * static_cast<size_t>(n_classes + 1)) + 1024;
}

int rls_query_lwork(int64_t d, int* lwork) {
  RlsLib* R;
  int rc;
  if ((rc = rls_lib(&R))) return rc;
  const int64_t P = gram_pitch(d);
  if (cusolverDnDpotrf_bufferSize(R->sol[0], CUBLAS_FILL_MODE_LOWER, static_cast<int>(d + 1), nullptr, static_cast<int>(P), lwork) !=
      CUSOLVER_STATUS_SUCCESS)
    return set_error(ODF_ERR_CUDA, "cusolverDnDpotrf_bufferSize failed");
  return ODF_OK;
}

// perm: rows sorted by class (stable); seg[c] .. seg[c+1] = the rows of class c in perm; row_class[r] = class of perm[r].
// Empty classes are skipped (their weights are left untouched).  SYNCHRONOUS (reads the factorisation status).
int rls_train(const float* X, int64_t n, int64_t d, int64_t ldx, const double* Yw, const int64_t* perm, const int64_t* seg_host,
              const int* row_class, int64_t n_classes, double lam, float* W, float* losses, void* ws, size_t ws_bytes,
              cudaStream_t st) {
  if (n <= 0 || d <= 0 || n_classes <= 0 || ldx < d || d + 5 > 0x7fffff00) return set_error(ODF_ERR_ARG, "rls_train: bad shape");
  RlsLib* R;
  int rc;
  if ((rc = rls_lib(&R))) return rc;
  int lwork = 0;
  if ((rc = rls_query_lwork(d, &lwork))) return rc;
  if (ws_bytes < rls_workspace_bytes(n, d, n_classes, lwork)) return set_error(ODF_ERR_WORKSPACE, "rls_train: workspace too small");
  const int64_t P = gram_pitch(d);
  uint8_t* p = static_cast<uint8_t*>(ws);
  auto take = [&](size_t bytes) { uint8_t* r = p; p += al256(bytes); return r; };
  double* G = reinterpret_cast<double*>(take(static_cast<size_t>(n_classes) * P * P * 8));
  double* Bm = reinterpret_cast<double*>(take(static_cast<size_t>(n_classes) * 4 * (d + 1) * 8));
  double* work[RLS_STREAMS];
  for (int s = 0; s < RLS_STREAMS; ++s) work[s] = reinterpret_cast<double*>(take(static_cast<size_t>(lwork > 0 ? lwork : 1) * 8));
  int* info = reinterpret_cast<int*>(take(sizeof(int) * static_cast<size_t>(2 * n_classes + 2)));
  int64_t* seg_dev = reinterpret_cast<int64_t*>(take(sizeof(int64_t) * static_cast<size_t>(n_classes + 1)));
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int) * static_cast<size_t>(2 * n_classes + 2), st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(seg_dev, seg_host, sizeof(int64_t) * static_cast<size_t>(n_classes + 1), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return set_cuda_error(e, "rls_train: memset / segment upload");
  // Gram matrices of every class, upper tiles
  const int n_tiles = static_cast<int>(P / GT);
  {
    dim3 grid(static_cast<unsigned>(n_tiles * (n_tiles + 1) / 2), 1, static_cast<unsigned>(n_classes));
    rls_gram_kernel<<<grid, 128, 0, st>>>(X, ldx, static_cast<int>(d), Yw, perm, seg_dev, n_tiles, P, G);
    if ((e = cudaGetLastError()) != cudaSuccess) return set_cuda_error(e, "rls_gram_kernel launch");
    dim3 g2(static_cast<unsigned>((d + 256) / 256), static_cast<unsigned>(n_classes));
    rls_rhs_kernel<<<g2, 256, 0, st>>>(G, P, static_cast<int>(d), lam, Bm);
    if ((e = cudaGetLastError()) != cudaSuccess) return set_cuda_error(e, "rls_rhs_kernel launch");
  }
  // factorise + solve, classes round-robin on the side streams (row-major upper == column-major lower)
  if ((e = cudaEventRecord(R->fork, st)) != cudaSuccess) return set_cuda_error(e, "rls_train: event");
  for (int s = 0; s < RLS_STREAMS; ++s)
    if ((e = cudaStreamWaitEvent(R->st[s], R->fork, 0)) != cudaSuccess) return set_cuda_error(e, "rls_train: fork");
  for (int64_t c = 0; c < n_classes; ++c) {
    if (seg_host[c + 1] == seg_host[c]) continue;
    const int s = static_cast<int>(c % RLS_STREAMS);
    double* Gc = G + c * P * P;
    if (cusolverDnDpotrf(R->sol[s], CUBLAS_FILL_MODE_LOWER, static_cast<int>(d + 1), Gc, static_cast<int>(P), work[s], lwork,
                         info + 2 * c) != CUSOLVER_STATUS_SUCCESS)
      return set_error(ODF_ERR_CUDA, "cusolverDnDpotrf (RLS) failed to launch");
    if (cusolverDnDpotrs(R->sol[s], CUBLAS_FILL_MODE_LOWER, static_cast<int>(d + 1), 4, Gc, static_cast<int>(P), Bm + c * 4 * (d + 1),
                         static_cast<int>(d + 1), info + 2 * c + 1) != CUSOLVER_STATUS_SUCCESS)
      return set_error(ODF_ERR_CUDA, "cusolverDnDpotrs (RLS) failed to launch");
  }
  for (int s = 0; s < RLS_STREAMS; ++s) {
    if ((e = cudaEventRecord(R->join[s], R->st[s])) != cudaSuccess) return set_cuda_error(e, "rls_train: event");
    if ((e = cudaStreamWaitEvent(st, R->join[s], 0)) != cudaSuccess) return set_cuda_error(e, "rls_train: join");
  }
  // losses and fp32 weights
  rls_loss_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, st>>>(X, ldx, static_cast<int>(d), Yw, perm, row_class, n, Bm, losses);
  if ((e = cudaGetLastError()) != cudaSuccess) return set_cuda_error(e, "rls_loss_kernel launch");
  const int64_t nw = n_classes * 4 * (d + 1);
  f64_to_f32_kernel<<<static_cast<unsigned>((nw + 255) / 256 > 1184 ? 1184 : (nw + 255) / 256), 256, 0, st>>>(Bm, W, nw);
  if ((e = cudaGetLastError()) != cudaSuccess) return set_cuda_error(e, "f64_to_f32_kernel launch");
  std::vector<int> hinfo(static_cast<size_t>(2 * n_classes), 0);
  e = cudaMemcpyAsync(hinfo.data(), info, sizeof(int) * hinfo.size(), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_cuda_error(e, "rls_train");
  for (size_t i = 0; i < hinfo.size(); ++i)
    if (hinfo[i] != 0) {
      char buf[160];
      snprintf(buf, sizeof buf, "RLS normal matrix of class %zu is not positive definite (info=%d)", i / 2, hinfo[i]);
      return set_error(ODF_ERR_LINALG, buf);
    }
  return ODF_OK;
}

int rls_apply(const float* feat, int64_t n, int64_t d, int64_t ldf, const float* Wp, const float* bias, const float* Tinv,
              const float* mu, const float* ex_boxes, int64_t C, float img_w, float img_h, float eps, const float* mean, float zscale,
              float* out, cudaStream_t st) {
  if (n <= 0) return ODF_OK;
  if (d <= 0 || C <= 0 || ldf < d) return set_error(ODF_ERR_ARG, "rls_apply: bad shape");
  const size_t smem = sizeof(float) * static_cast<size_t>(d + 4 * C);
  if (smem > 200 * 1024) return set_error(ODF_ERR_ARG, "rls_apply: feature row too long for shared memory");
  static DeviceOnce once;
  bool& set = once.here();
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(rls_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(rls_apply_kernel)");
    set = true;
  }
  rls_apply_kernel<<<static_cast<unsigned>(n), 128, smem, st>>>(feat, ldf, static_cast<int>(d), Wp, bias, Tinv, mu, ex_boxes,
                                                                static_cast<int>(C), img_w, img_h, eps, mean, zscale, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ODF_OK : set_cuda_error(e, "rls_apply_kernel launch");
}

}  // namespace odf
