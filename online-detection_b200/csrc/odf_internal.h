// Internal declarations shared by the translation units of libodf (not part of the C ABI).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include "../../include/odf.h"

struct CUtensorMap_st;   // <cuda.h>

namespace odf {

enum : int { MODE_MMV = 0, MODE_STORE = 1 };
enum : int { KIND_TF32 = ODF_KIND_TF32, KIND_F16 = ODF_KIND_F16 };

// Kernel-side parameters of the fused Gaussian tile (see odf_gauss_tile.cu).
struct TileParams {
  int n_rows, n_cols;      // points on the row / column side
  int kblocks;             // 128-byte k-blocks of real features (d_pad / 32 tf32 or d_pad / 64 fp16)
  int T_pad;               // padded number of right-hand sides (16 or 32)
  int mode;                // MODE_MMV or MODE_STORE
  int n_rowblocks, n_coltiles;
  int n_splits, tiles_per_split, group_rows;
  int store_vec4;
  float neg_scale_log2;    // -log2(e) / (2 sigma^2)
  const float* r_scale;    // device scalars: power-of-two scaling of the row / column point set
  const float* q_scale;
  const float* rnorm;      // |row point|^2, padded to a multiple of 128 entries
  const float* qnorm;      // |column point|^2, padded to a multiple of 128 entries
  float* out;              // MODE_MMV: partial slabs [n_splits][n_rows][T_pad]; MODE_STORE: K
  int64_t ldo;             // MODE_STORE: row pitch of K
  int64_t split_stride;    // MODE_MMV: elements between split slabs
  float* panel;            // optional (MODE_MMV): spill K tiles here, [n_rows x ldpanel] fp32
  int64_t ldpanel;
  const uint32_t* v_absmax; // pair kernel: 32 words fixing the per-column scales of the fp16 V^T split
  __half* panel16;         // optional (MODE_MMV, SPILL16 kernels): fp16 hi / lo planes of the K tiles, tile-blocked
  int64_t panel16_plane;   // elements between the hi and the lo plane (n_coltiles * n_rowblocks * 128 * 128)
  int dbg;                 // bring-up timing experiments (env ODF_TILE_DEBUG; results are garbage when set):
                           // 1 = no TMA refills, 2 = no S MMAs, 4 = no epilogue math, 8 = no K.V MMAs, 16 = print clocks
  // LINEAR store variant only (appended: the layout seen by the other instantiations is unchanged):
  // out = lin_alpha * (x . c) + lin_beta * out, operands prepared with a zero seed block
  float lin_alpha, lin_beta;
};

// Host-side launch description.
struct TileLaunch {
  int kind;                    // KIND_TF32 or KIND_F16
  const void *r_hi, *r_lo;
  const float *r_norm, *r_scale;
  int64_t n_rows;
  const void *q_hi, *q_lo;
  const float *q_norm, *q_scale;
  int64_t n_cols;
  int64_t d_pad;               // padded feature count (multiple of the k-block width)
  const float *vt_hi, *vt_lo;  // [T_pad x ldvt], zero beyond n_cols
  int64_t ldvt;
  int T_pad;
  int mode;
  int n_splits;
  float sigma;
  float* out;
  int64_t ldo;
  int64_t split_stride;
  float* panel;                // optional K spill (MODE_MMV)
  int64_t ldpanel;
  void* panel16;               // optional fp16-plane K spill for odf_panel16_tmm (see odf_panel16.cu)
  const void *vt16_hi, *vt16_lo;   // optional fp16 split of V^T [T_pad x ldvt16] (odf_split_rhs16): enables the pair kernel
  int64_t ldvt16;
  const uint32_t* v_absmax;
  int linear;                  // MODE_STORE only: store lin_alpha * (x . c) + lin_beta * out instead of the Gaussian kernel
  float lin_alpha, lin_beta;
};

int launch_gauss_tile(const TileLaunch& L, cudaStream_t stream);
// CTA-pair (cta_group::2) variant for large MODE_MMV launches (odf_gauss_tile2.cu)
bool tile2_eligible(const TileLaunch& L);
bool tile2_rows_eligible(int64_t n_rows);
int split_rhs16(const float* V, int64_t m, int64_t T, int64_t ldv, float scale, uint32_t* absmax, void* vt_hi, void* vt_lo,
                int64_t ldvt, int T_pad, cudaStream_t st);
int launch_gauss_tile2(const TileLaunch& L, cudaStream_t stream);
int make_map_plain_f32(::CUtensorMap_st* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                       int box_cols);
int make_map_sw128(::CUtensorMap_st* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int esize);
int make_map_plain_f16(::CUtensorMap_st* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols);
int make_map_plain_u8(::CUtensorMap_st* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols);
int panel_splits(int64_t n_rows, int64_t M);
// fp16-plane panel (tensor-core contraction, odf_panel16.cu)
size_t panel16_bytes(int64_t n_rows, int64_t M);
int panel16_splits(int64_t n_rows, int64_t M);
int launch_panel16_tmm(const void* P16, int64_t n_rows, int64_t M, const void* W16, const uint32_t* absmax, int T_pad,
                       int n_splits, float* out_partial, cudaStream_t st, int hi_only = 0);
int panel16_mmv_splits(int64_t n_rows, int64_t M);
int launch_panel16_mmv(const void* P16, int64_t n_rows, int64_t M, const void* V16, const uint32_t* absmax, int T_pad,
                       int n_splits, float* out_partial, cudaStream_t st, int hi_only = 0);
int finish_w16(const float* partial, int S, int64_t n, int T_pad, int64_t T, const float* addend, int64_t ld_add,
               float* Wf, uint32_t* absmax, void* W16, cudaStream_t st);
// lo plane of the K panel: the residual K - rn16(K) (|.| <= 2^-12 for K <= 1) in fixed point, one byte:
// u = clamp(rni(r 2^19), -128, 127) + 128   (2^-20 absolute; widened back to fp16 by widen_lo8 in odf_panel16_common.cuh)
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lo8_of(float r) {
  const int i = max(-128, min(127, __float2int_rn(r * 524288.f)));
  return static_cast<uint32_t>(i + 128);
}
#endif
// power-of-two scaling of W for the fp16 split: s * max|W| in [2^14, 2^15)
__host__ __device__ inline float w16_scale_from_bits(uint32_t absmax_bits, bool inverse) {
  const int e = static_cast<int>((absmax_bits >> 23) & 0xffu);
  if (e == 0 || e == 255) return 1.f;
  int se = 127 + 14 - (e - 127);
  if (se < 1) se = 1;
  if (se > 253) se = 253;
  if (inverse) se = 254 - se;
  const uint32_t bits = static_cast<uint32_t>(se) << 23;
#ifdef __CUDA_ARCH__
  return __uint_as_float(bits);
#else
  float f;
  memcpy(&f, &bits, 4);
  return f;
#endif
}
int launch_panel_tmm(const float* P, int64_t ldp, const float* W, int64_t n_rows, int64_t M, int T_pad,
                     int n_splits, float* out_partial, cudaStream_t st);
int tile_default_splits(int64_t n_rows, int64_t n_cols, int64_t row_bytes);

// error plumbing (thread-local last-error string behind odf_last_error())
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);


// Per-device one-shot state (cudaFuncSetAttribute is per device; a process may touch several GPUs).
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= kMaxDevices) ? 0 : dev;
}
struct DeviceOnce {
  bool done[kMaxDevices] = {};
  bool& here() { return done[current_device()]; }
};
inline int device_sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

inline int64_t kblock_elems(int kind) { return kind == KIND_F16 ? 64 : 32; }
// operand row pitch in elements: padded features + seed block + ones block
inline int64_t operand_pitch(int64_t d, int kind) {
  const int64_t bk = kblock_elems(kind);
  return (d + bk - 1) / bk * bk + 2 * bk;
}
inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace odf
