// Internal declarations shared by the translation units of libodf (not part of the C ABI).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

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
  int dbg;                 // bring-up timing experiments (env ODF_TILE_DEBUG; results are garbage when set):
                           // 1 = no TMA refills, 2 = no S MMAs, 4 = no epilogue math, 8 = no K.V MMAs, 16 = print clocks
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
};

int launch_gauss_tile(const TileLaunch& L, cudaStream_t stream);
int make_map_plain_f32(::CUtensorMap_st* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                       int box_cols);
int panel_splits(int64_t n_rows, int64_t M);
int launch_panel_tmm(const float* P, int64_t ldp, const float* W, int64_t n_rows, int64_t M, int T_pad,
                     int n_splits, float* out_partial, cudaStream_t st);
int tile_default_splits(int64_t n_rows, int64_t n_cols, int64_t row_bytes);

// error plumbing (thread-local last-error string behind odf_last_error())
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);

inline int64_t kblock_elems(int kind) { return kind == KIND_F16 ? 64 : 32; }
// operand row pitch in elements: padded features + seed block + ones block
inline int64_t operand_pitch(int64_t d, int kind) {
  const int64_t bk = kblock_elems(kind);
  return (d + bk - 1) / bk * bk + 2 * bk;
}
inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace odf
