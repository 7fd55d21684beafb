// Integer / index side of the on-line detection path, on the GPU:
//
//  * odf_select_indices + odf_gather_rows: the minibootstrap's hard- / easy-negative selection,
//    `torch.where(scores > thr)[0]` followed by `X[idx]` / `torch.cat`
//    (src/modules/region-classifier/OnlineRegionClassifier_incore.py:117-137), as a STABLE stream
//    compaction (ascending original index, bit-identical to torch.where) and a row gather whose row count
//    is read on the device, so the selected rows land in a pre-allocated cache without host round trips.
//  * odf_decode_boxes: py_od_utils.decode_boxes_detector (src/py_od_utils.py:247-274), legacy "+1" box
//    convention, clamped to the image.
//  * odf_detect_postprocess: OnlineDetectionPostProcessor.filter_results
//    (src/modules/accuracy-evaluator/OnlineDetectionPostProcessor.py:35-79): `score > thresh` (strict),
//    per-class greedy NMS in descending score order with IoU(+1) > nms_thresh (strict) -- the semantics of
//    maskrcnn-benchmark's `_C.nms` / boxlist_nms, kept indices in ascending original order -- classes
//    concatenated 1..T, then the `kthvalue` top-K rule (`score >= K-th largest`, ties all kept).
//
// Everything here is index work on a few thousand boxes per image: the kernels are latency-sized (one CTA
// per class, bit-mask suppression matrix, warp-wide sequential sweep), deterministic, and run on the
// caller's stream with no host synchronisation; counts stay on the device.
#include "odf_internal.h"

namespace odf {
namespace {

constexpr int SEL_BLOCK = 1024;

// ------------------------------------------------------------------------------------------- compaction
__device__ __forceinline__ bool pred_of(float s, float thr, int strict) { return strict ? (s > thr) : (s >= thr); }

__global__ void __launch_bounds__(SEL_BLOCK)
select_count_kernel(const float* __restrict__ scores, int64_t n, int64_t stride, float thr, int strict, int* block_counts) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SEL_BLOCK + threadIdx.x;
  const bool p = i < n && pred_of(scores[i * stride], thr, strict);
  const int c = __syncthreads_count(p);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// single CTA: exclusive scan of the block counts (in place) + total
__global__ void __launch_bounds__(1024)
select_scan_kernel(int* block_counts, int n_blocks, int* total) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? block_counts[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_sums[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_sums[threadIdx.x] = w;
    }
    __syncthreads();
    const int prefix = carry + (threadIdx.x >= 32 ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + x - v;
    if (i < n_blocks) block_counts[i] = prefix;
    __syncthreads();
    if (threadIdx.x == 1023) carry = prefix + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SEL_BLOCK)
select_scatter_kernel(const float* __restrict__ scores, int64_t n, int64_t stride, float thr, int strict,
                      const int* __restrict__ block_offsets, int64_t* __restrict__ idx_out) {
  __shared__ int warp_sums[32];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * SEL_BLOCK + threadIdx.x;
  const bool p = i < n && pred_of(scores[i * stride], thr, strict);
  const unsigned ballot = __ballot_sync(0xffffffffu, p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_sums[warp] = __popc(ballot);
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    warp_sums[lane] = w;      // inclusive
  }
  __syncthreads();
  if (p) {
    const int pos = block_offsets[blockIdx.x] + (warp ? warp_sums[warp - 1] : 0) + __popc(ballot & ((1u << lane) - 1u));
    idx_out[pos] = i;
  }
}

// dst[k, :] = src[idx[k], :] for k < *count  (one warp per row, float4 when aligned)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, int64_t ld_src, const int64_t* __restrict__ idx,
                   const int* __restrict__ count, int64_t d, float* __restrict__ dst, int64_t ld_dst, int vec4) {
  const int n = *count;
  const int lane = threadIdx.x & 31;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); k < n; k += static_cast<int64_t>(gridDim.x) * 8) {
    const float* s = src + idx[k] * ld_src;
    float* t = dst + k * ld_dst;
    if (vec4) {
      for (int64_t c = lane; c < d / 4; c += 32) reinterpret_cast<float4*>(t)[c] = __ldg(reinterpret_cast<const float4*>(s) + c);
    } else {
      for (int64_t c = lane; c < d; c += 32) t[c] = __ldg(s + c);
    }
  }
}

// ------------------------------------------------------------------------------------------- box decode
__global__ void decode_boxes_kernel(const float* __restrict__ ex, const float* __restrict__ deltas, int R, int Tc,
                                    float img_w, float img_h, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * Tc) return;
  const int r = i / Tc;
  const float4 e = __ldg(reinterpret_cast<const float4*>(ex) + r);
  const float4 dl = __ldg(reinterpret_cast<const float4*>(deltas) + i);
  // same operation order as the reference (fp32, no fused multiply-add across its statements)
  const float w = __fadd_rn(__fsub_rn(e.z, e.x), 1.f), h = __fadd_rn(__fsub_rn(e.w, e.y), 1.f);
  const float cx = __fadd_rn(e.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(e.y, __fmul_rn(0.5f, h));
  const float pcx = __fadd_rn(__fmul_rn(dl.x, w), cx), pcy = __fadd_rn(__fmul_rn(dl.y, h), cy);
  const float pw = __fmul_rn(expf(dl.z), w), ph = __fmul_rn(expf(dl.w), h);
  float4 o;
  o.x = fmaxf(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), 0.f);
  o.y = fmaxf(__fsub_rn(pcy, __fmul_rn(0.5f, ph)), 0.f);
  o.z = fminf(__fsub_rn(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 1.f), img_w - 1.f);
  o.w = fminf(__fsub_rn(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), 1.f), img_h - 1.f);
  reinterpret_cast<float4*>(out)[i] = o;
}

// ------------------------------------------------------------------------------------------- NMS
// IoU with the legacy +1 convention, the arithmetic of maskrcnn-benchmark's nms kernels (devIoU):
// inter / (areaA + areaB - inter) in fp32.
__device__ __forceinline__ float iou_plus1(const float4 a, const float4 b) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f), h = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
  const float sb = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
}

// One CTA per class j (1..Tc-1): candidates (score > thresh), stable descending rank, sorted boxes.
//   order[j][rank] = original RoI index; n_cand[j]
__global__ void __launch_bounds__(1024)
nms_rank_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int R, int Tc, float score_thresh,
                int* __restrict__ order, float4* __restrict__ sorted_boxes, int* __restrict__ n_cand) {
  extern __shared__ float s_sc[];                 // R scores of this class (NaN-free assumed)
  const int j = blockIdx.x + 1;
  for (int i = threadIdx.x; i < R; i += blockDim.x) s_sc[i] = scores[static_cast<int64_t>(i) * Tc + j];
  __syncthreads();
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    const float si = s_sc[i];
    if (si > score_thresh) {
      int rank = 0;
      for (int k = 0; k < R; ++k) {
        const float sk = s_sc[k];
        rank += (sk > score_thresh) && (sk > si || (sk == si && k < i));
      }
      order[static_cast<int64_t>(j) * R + rank] = i;
      sorted_boxes[static_cast<int64_t>(j) * R + rank] = __ldg(reinterpret_cast<const float4*>(boxes) + static_cast<int64_t>(i) * Tc + j);
    }
  }
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  int mine = 0;
  for (int i = threadIdx.x; i < R; i += blockDim.x) mine += s_sc[i] > score_thresh;
  atomicAdd(&total, mine);
  __syncthreads();
  if (threadIdx.x == 0) n_cand[j] = total;
}

// mask[j][i][w] bit b: sorted box i suppresses sorted box (64 w + b) (only later boxes)
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ sorted_boxes, const int* __restrict__ n_cand, int R, int words,
                float nms_thresh, unsigned long long* __restrict__ mask) {
  const int j = blockIdx.z + 1;
  const int n = n_cand[j];
  const int row0 = blockIdx.y * 64, col0 = blockIdx.x * 64;
  if (row0 >= n || col0 >= n || col0 + 63 < row0) return;     // nothing, or entirely at/below the diagonal
  __shared__ float4 cb[64];
  const float4* sb = sorted_boxes + static_cast<int64_t>(j) * R;
  if (col0 + threadIdx.x < n) cb[threadIdx.x] = sb[col0 + threadIdx.x];
  __syncthreads();
  const int i = row0 + threadIdx.x;
  if (i < n) {
    const float4 a = sb[i];
    unsigned long long bits = 0;
    const int lim = min(64, n - col0);
    for (int b = (col0 == row0 ? threadIdx.x + 1 : 0); b < lim; ++b)
      if (col0 + b > i && iou_plus1(a, cb[b]) > nms_thresh) bits |= 1ull << b;
    mask[(static_cast<int64_t>(j) * R + i) * words + blockIdx.x] = bits;
  }
}

// one warp per class: sequential greedy sweep over the sorted candidates; keep[r*Tc + j] = 1 for kept RoIs
__global__ void __launch_bounds__(32)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ order, const int* __restrict__ n_cand,
                 int R, int Tc, int words, unsigned char* __restrict__ keep) {
  const int j = blockIdx.x + 1;
  const int n = n_cand[j];
  const int lane = threadIdx.x;
  // lane l owns suppression words l, l+32, ... (R <= 64 * 32 * 4)
  unsigned long long remv[4] = {0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    const int w = i >> 6;
    unsigned long long mine = remv[0];
#pragma unroll
    for (int q = 1; q < 4; ++q) mine = ((w >> 5) == q) ? remv[q] : mine;
    const unsigned long long word = __shfl_sync(0xffffffffu, mine, w & 31);
    if (!((word >> (i & 63)) & 1ull)) {
      if (lane == 0) keep[static_cast<int64_t>(order[static_cast<int64_t>(j) * R + i]) * Tc + j] = 1;
      const unsigned long long* mrow = mask + (static_cast<int64_t>(j) * R + i) * words;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ww = lane + 32 * q;
        if (ww < words && ww >= w) remv[q] |= mrow[ww];
      }
    }
  }
}

// single CTA: class-major, index-ascending compaction of the keep flags -> detections + count
__global__ void __launch_bounds__(1024)
nms_collect_kernel(const unsigned char* __restrict__ keep, const float* __restrict__ boxes, const float* __restrict__ scores,
                   int R, int Tc, float* __restrict__ out_boxes, float* __restrict__ out_scores,
                   int64_t* __restrict__ out_labels, int64_t* __restrict__ out_rois, int* __restrict__ count) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int total = R * (Tc - 1);
  for (int base = 0; base < total; base += 1024) {
    const int e = base + threadIdx.x;              // e = (j-1) * R + r
    int j = 0, r = 0;
    bool p = false;
    if (e < total) {
      j = e / R + 1;
      r = e - (j - 1) * R;
      p = keep[static_cast<int64_t>(r) * Tc + j] != 0;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, p);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_sums[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    if (p) {
      const int pos = carry + (warp ? warp_sums[warp - 1] : 0) + __popc(ballot & ((1u << lane) - 1u));
      reinterpret_cast<float4*>(out_boxes)[pos] = __ldg(reinterpret_cast<const float4*>(boxes) + static_cast<int64_t>(r) * Tc + j);
      out_scores[pos] = scores[static_cast<int64_t>(r) * Tc + j];
      out_labels[pos] = j;
      out_rois[pos] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_sums[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

// top-K rule of filter_results: if count > K keep the detections whose score is >= the K-th largest
// (kthvalue(n - K + 1)); ties at the threshold are all kept.  flag[i] = 1 if detection i survives.
__global__ void __launch_bounds__(256)
topk_flag_kernel(const float* __restrict__ det_scores, const int* __restrict__ count, int K, unsigned char* __restrict__ flag) {
  const int n = *count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    unsigned char f = 1;
    if (K > 0 && n > K) {
      const float si = det_scores[i];
      int greater = 0;
      for (int k = 0; k < n; ++k) greater += det_scores[k] > si;
      f = greater < K;
    }
    flag[i] = f;
  }
}

// single CTA: in-order compaction of the flagged detections (src -> dst buffers), final count
__global__ void __launch_bounds__(1024)
topk_compact_kernel(const unsigned char* __restrict__ flag, const int* __restrict__ count_in,
                    const float* __restrict__ b_in, const float* __restrict__ s_in, const int64_t* __restrict__ l_in,
                    const int64_t* __restrict__ r_in, float* __restrict__ b_out, float* __restrict__ s_out,
                    int64_t* __restrict__ l_out, int64_t* __restrict__ r_out, int* __restrict__ count_out) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int n = *count_in;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const bool p = i < n && flag[i] != 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, p);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_sums[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    if (p) {
      const int pos = carry + (warp ? warp_sums[warp - 1] : 0) + __popc(ballot & ((1u << lane) - 1u));
      reinterpret_cast<float4*>(b_out)[pos] = reinterpret_cast<const float4*>(b_in)[i];
      s_out[pos] = s_in[i];
      l_out[pos] = l_in[i];
      r_out[pos] = r_in[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_sums[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *count_out = carry;
}

inline size_t al256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }

}  // namespace
}  // namespace odf

using namespace odf;

extern "C" {

size_t odf_select_workspace_bytes(int64_t n) {
  return al256(sizeof(int) * static_cast<size_t>((n + SEL_BLOCK - 1) / SEL_BLOCK + 1));
}

int odf_select_indices(const float* scores, int64_t n, int64_t stride, float thresh, int strict, int64_t* idx_out,
                       int* count_out, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n < 0 || stride < 1) return set_error(ODF_ERR_ARG, "select_indices: bad shape");
  if (n == 0) {
    cudaError_t e = cudaMemsetAsync(count_out, 0, sizeof(int), st);
    return e == cudaSuccess ? ODF_OK : set_cuda_error(e, "select_indices");
  }
  if (ws == nullptr || ws_bytes < odf_select_workspace_bytes(n)) return set_error(ODF_ERR_WORKSPACE, "select_indices: workspace too small");
  const int n_blocks = static_cast<int>((n + SEL_BLOCK - 1) / SEL_BLOCK);
  int* bc = static_cast<int*>(ws);
  select_count_kernel<<<n_blocks, SEL_BLOCK, 0, st>>>(scores, n, stride, thresh, strict, bc);
  select_scan_kernel<<<1, 1024, 0, st>>>(bc, n_blocks, count_out);
  select_scatter_kernel<<<n_blocks, SEL_BLOCK, 0, st>>>(scores, n, stride, thresh, strict, bc, idx_out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ODF_OK : set_cuda_error(e, "select_indices launch");
}

int odf_gather_rows(const float* src, int64_t ld_src, const int64_t* idx, const int* count, int64_t max_rows, int64_t d,
                    float* dst, int64_t ld_dst, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (max_rows <= 0 || d <= 0) return ODF_OK;
  const int vec4 = (d % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? 1 : 0;
  int64_t blocks = (max_rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(src, ld_src, idx, count, d, dst, ld_dst, vec4);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ODF_OK : set_cuda_error(e, "gather_rows launch");
}

int odf_decode_boxes(const float* ex_boxes, const float* deltas, int64_t R, int64_t Tc, float img_w, float img_h, float* out,
                     void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (R <= 0 || Tc <= 0) return ODF_OK;
  if ((reinterpret_cast<uintptr_t>(ex_boxes) & 15) || (reinterpret_cast<uintptr_t>(deltas) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(ODF_ERR_ARG, "decode_boxes: buffers must be 16-byte aligned");
  const int64_t total = R * Tc;
  decode_boxes_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(ex_boxes, deltas, static_cast<int>(R),
                                                                                static_cast<int>(Tc), img_w, img_h, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ODF_OK : set_cuda_error(e, "decode_boxes launch");
}

size_t odf_postprocess_workspace_bytes(int64_t R, int64_t Tc) {
  const size_t words = static_cast<size_t>((R + 63) / 64);
  const size_t cap = static_cast<size_t>(R) * static_cast<size_t>(Tc);
  return al256(sizeof(int) * cap) + al256(sizeof(float4) * cap) + al256(sizeof(int) * (Tc + 1)) +
         al256(sizeof(unsigned long long) * cap * words) + al256(cap) +                         /* order, sorted, n_cand, mask, keep */
         al256(sizeof(float4) * cap) + al256(sizeof(float) * cap) + 2 * al256(sizeof(int64_t) * cap) +  /* staged detections */
         al256(cap) + al256(sizeof(int) * 2);                                                     /* top-k flags, counts */
}

int odf_detect_postprocess(const float* boxes, const float* scores, int64_t R, int64_t Tc, float score_thresh,
                           float nms_thresh, int dets_per_img, float* out_boxes, float* out_scores, int64_t* out_labels,
                           int64_t* out_rois, int* out_count, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (R < 0 || Tc < 2) return set_error(ODF_ERR_ARG, "detect_postprocess: need R >= 0 and at least one foreground class");
  if (R > 8192) return set_error(ODF_ERR_ARG, "detect_postprocess: at most 8192 RoIs per image");
  if ((reinterpret_cast<uintptr_t>(boxes) & 15) || (reinterpret_cast<uintptr_t>(out_boxes) & 15))
    return set_error(ODF_ERR_ARG, "detect_postprocess: box buffers must be 16-byte aligned");
  if (R == 0) {
    cudaError_t e0 = cudaMemsetAsync(out_count, 0, sizeof(int), st);
    return e0 == cudaSuccess ? ODF_OK : set_cuda_error(e0, "detect_postprocess");
  }
  if (ws == nullptr || ws_bytes < odf_postprocess_workspace_bytes(R, Tc)) return set_error(ODF_ERR_WORKSPACE, "detect_postprocess: workspace too small");
  const int Ri = static_cast<int>(R), Tci = static_cast<int>(Tc);
  const int words = (Ri + 63) / 64;
  const size_t cap = static_cast<size_t>(R) * static_cast<size_t>(Tc);
  uint8_t* p = static_cast<uint8_t*>(ws);
  auto take = [&](size_t bytes) { uint8_t* r = p; p += al256(bytes); return r; };
  int* order = reinterpret_cast<int*>(take(sizeof(int) * cap));
  float4* sorted = reinterpret_cast<float4*>(take(sizeof(float4) * cap));
  int* n_cand = reinterpret_cast<int*>(take(sizeof(int) * (Tc + 1)));
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * cap * words));
  unsigned char* keep = take(cap);
  float* sb = reinterpret_cast<float*>(take(sizeof(float4) * cap));
  float* ss = reinterpret_cast<float*>(take(sizeof(float) * cap));
  int64_t* sl = reinterpret_cast<int64_t*>(take(sizeof(int64_t) * cap));
  int64_t* sr = reinterpret_cast<int64_t*>(take(sizeof(int64_t) * cap));
  unsigned char* flag = take(cap);
  int* cnt = reinterpret_cast<int*>(take(sizeof(int) * 2));
  cudaError_t e = cudaMemsetAsync(keep, 0, cap, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(mask, 0, sizeof(unsigned long long) * cap * words, st);
  if (e != cudaSuccess) return set_cuda_error(e, "detect_postprocess memset");
  const int nt = Ri < 1024 ? ((Ri + 31) / 32 * 32) : 1024;
  nms_rank_kernel<<<Tci - 1, nt, sizeof(float) * Ri, st>>>(boxes, scores, Ri, Tci, score_thresh, order, sorted, n_cand);
  dim3 mg(words, words, Tci - 1);
  nms_mask_kernel<<<mg, 64, 0, st>>>(sorted, n_cand, Ri, words, nms_thresh, mask);
  nms_sweep_kernel<<<Tci - 1, 32, 0, st>>>(mask, order, n_cand, Ri, Tci, words, keep);
  nms_collect_kernel<<<1, 1024, 0, st>>>(keep, boxes, scores, Ri, Tci, sb, ss, sl, sr, cnt);
  topk_flag_kernel<<<64, 256, 0, st>>>(ss, cnt, dets_per_img, flag);
  topk_compact_kernel<<<1, 1024, 0, st>>>(flag, cnt, sb, ss, sl, sr, out_boxes, out_scores, out_labels, out_rois, out_count);
  e = cudaGetLastError();
  return e == cudaSuccess ? ODF_OK : set_cuda_error(e, "detect_postprocess launch");
}

}  // extern "C"
