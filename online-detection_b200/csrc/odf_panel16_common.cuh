// Shared-memory stage geometry, UMMA descriptors and copy helpers common to the fp16-plane panel kernels
// (odf_panel16.cu: K^T W and K V as separate passes; odf_panel16_sweep.cu: both in one pass over L2-sized row groups).
#pragma once
#include <cuda_fp16.h>
#include "odf_ptx.cuh"

namespace odf {
namespace p16 {

constexpr int QR = 64;                    // panel rows (MMA K dimension) per pipeline stage
constexpr int QCHUNK = QR * 128;          // 8 KB: [64 rows x 128 B] = one MN-major SWIZZLE_128B chunk (64 MN values)
constexpr int QSTAGE = 5 * QCHUNK;        // P_hi h0, P_hi h1, P_lo h0, P_lo h1, W16
constexpr int QNS = 5;
constexpr int QBOX = 32 * 128 * 2;        // 8 KB: one TMA box of P = [16 centre groups][32 rows][8 fp16]
constexpr int QFLUSH = 8;                 // stages per TMEM accumulation chain (512 rows)
constexpr int QBARS = 2 * QNS + 4;
constexpr int QSMEM = QNS * QSTAGE + QBARS * 8 + 16 + 1024;
constexpr uint32_t QTM_COLS = 256;        // two accumulator buffers x (acc1 64 + acc2 64) columns

// B: MN-major SWIZZLE_128B operand: 64-value chunks along N are QCHUNK bytes apart (LBO), 8-row groups along K are
// 1024 B apart (SBO)  — validated with tools/mn_probe.cu.
constexpr uint64_t kSdescMnHi = (static_cast<uint64_t>(QCHUNK >> 4) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                                (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
// A: MN-major, no swizzle: core matrices [8 rows x 16 B] of 128 contiguous bytes; the next core matrix along K (rows)
// is 128 B further (LBO), the next group of 8 centres 32 rows x 16 B = 512 B further (SBO).
constexpr uint64_t kSdescMnPlainHi = (static_cast<uint64_t>(128 >> 4) << 16) | (static_cast<uint64_t>(512 >> 4) << 32) |
                                     (static_cast<uint64_t>(1) << 46);
constexpr uint64_t kSdescMnPlainHiSwapped = (static_cast<uint64_t>(512 >> 4) << 16) | (static_cast<uint64_t>(128 >> 4) << 32) |
                                            (static_cast<uint64_t>(1) << 46);
// kind::f16, fp16 A/B, fp32 accumulate, A and B MN-major (bits 15, 16), N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// ---- K . V orientation (rows are the MMA's M dimension, centres its K dimension) ----
constexpr int VA = 8 * 2048;              // [8 centre groups][128 rows][8 fp16] of one plane
constexpr int VSTAGE = 2 * VA + QCHUNK;   // P_hi, P_lo, V16 (64 centres x 128 B)  = 40 KB
static_assert(VSTAGE == QSTAGE, "both panel kernels use the same shared-memory budget");
// A: K-major, no swizzle: core matrices [8 rows x 16 B]; next core matrix along K (centres) 2048 B (LBO), next 8 rows 128 B (SBO)
constexpr uint64_t kSdescKPlainHi = (static_cast<uint64_t>(2048 >> 4) << 16) | (static_cast<uint64_t>(128 >> 4) << 32) |
                                    (static_cast<uint64_t>(1) << 46);
constexpr uint64_t kSdescKPlainHiSwapped = (static_cast<uint64_t>(128 >> 4) << 16) | (static_cast<uint64_t>(2048 >> 4) << 32) |
                                           (static_cast<uint64_t>(1) << 46);
// kind::f16, fp16 A/B, fp32 accumulate, A K-major, B MN-major (bit 16), N = 64, M = 128
constexpr uint32_t kIdescV = (1u << 4) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

}  // namespace p16
}  // namespace odf
