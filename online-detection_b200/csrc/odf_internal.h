// Internal declarations shared by the translation units of libodf (not part of the C ABI).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/odf.h"

namespace odf {

enum : int { MODE_MMV = 0, MODE_STORE = 1 };

// Kernel-side parameters of the fused Gaussian tile (see odf_gauss_tile.cu).
struct TileParams {
  int n_rows, n_cols;      // points on the row / column side
  int kblocks;             // d_pad / 32
  int T_pad;               // padded number of right-hand sides (16 or 32)
  int mode;                // MODE_MMV or MODE_STORE
  int n_rowblocks, n_coltiles;
  int n_splits, tiles_per_split, group_rows;
  int store_vec4;
  float neg_scale_log2;    // -log2(e) / (2 sigma^2)
  const float* rnorm;      // |row point|^2, padded to a multiple of 128 entries
  const float* qnorm;      // |column point|^2, padded to a multiple of 128 entries
  float* out;              // MODE_MMV: partial slabs [n_splits][n_rows][T_pad]; MODE_STORE: K
  int64_t ldo;             // MODE_STORE: row pitch of K
  int64_t split_stride;    // MODE_MMV: elements between split slabs
};

// Host-side launch description.
struct TileLaunch {
  const float *r_hi, *r_lo, *r_norm;
  int64_t n_rows;
  const float *q_hi, *q_lo, *q_norm;
  int64_t n_cols;
  int64_t d_pad;
  const float *vt_hi, *vt_lo;  // [T_pad x ldvt], zero beyond n_cols
  int64_t ldvt;
  int T_pad;
  int mode;
  int n_splits;
  float sigma;
  float* out;
  int64_t ldo;
  int64_t split_stride;
};

int launch_gauss_tile(const TileLaunch& L, cudaStream_t stream);
int tile_default_splits(int64_t n_rows, int64_t n_cols, int64_t d_pad);

// error plumbing (thread-local last-error string behind odf_last_error())
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace odf
