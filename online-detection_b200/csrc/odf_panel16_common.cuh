// Shared-memory stage geometry, UMMA descriptors and copy helpers common to the fp16-plane panel kernels
// (odf_panel16.cu: K^T W and K V as separate passes; odf_panel16_sweep.cu: both in one pass over L2-sized row groups).
#pragma once
#include <cuda_fp16.h>
#include "odf_ptx.cuh"

namespace odf {
namespace p16 {

constexpr int QR = 64;                    // panel rows (MMA K dimension) per pipeline stage
constexpr int QCHUNK = QR * 128;          // 8 KB: [64 rows x 128 B] = one MN-major SWIZZLE_128B chunk (64 MN values)
constexpr int QSTAGE = 5 * QCHUNK;        // P_hi h0, P_hi h1, P_lo h0, P_lo h1 (expanded in place from 8 KB of int8), W16
constexpr int QNS = 5;
constexpr int QBOX = 32 * 128 * 2;        // 8 KB: one TMA box of P = [16 centre groups][32 rows][8 fp16]
constexpr int QFLUSH = 8;                 // stages per TMEM accumulation chain (512 rows)
constexpr int QBARS = 3 * QNS + 4;        // FULL, EMPTY, accumulator FULL / EMPTY x 2, CONVERTED
constexpr int QTHREADS = 384;             // producer, MMA issuer, TMEM allocator, (idle), 4 epilogue warps, 4 converter warps
constexpr float kLoScale = 1.f / 524288.f;   // the lo plane holds rni((K - hi) 2^19) + 128 as bytes
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
// the same with N = 32: the lo plane is contracted with the hi half of W16 / V16 only (lo . lo is dropped anyway)
constexpr uint32_t kIdescN32 = (1u << 4) | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

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
constexpr uint32_t kIdescVN32 = (1u << 4) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

// The lo plane travels as one byte per kernel value, u = rni((K - hi) 2^19) + 128 (fixed point: 2^-20 absolute on K <= 1):
// a stage's 8 KB of bytes land in the UPPER half of its 16 KB fp16 lo area and the 4 converter warps widen them in place
// to the fp16 integers u - 128 the tensor core contracts (byte i -> half i, so the operand layout is the hi plane's).
// fp16(1024 + u) has the bit pattern 0x6400 | u; subtracting 1152 leaves u - 128 exactly.  Every thread reads its 64
// bytes, the group barrier separates all reads from all writes (the areas overlap), then its 128 bytes are written.
__device__ __forceinline__ void widen_lo8(uint8_t* lo_area, int ctid) {
  // thread t takes the 8-byte chunks t, t + 128, ...: consecutive lanes read consecutive 8 bytes and write consecutive 16 bytes
  // (a first version gave every thread 64 contiguous bytes: 16- / 32-way bank conflicts made the converters, not HBM, the
  // bound of the kernel: 5.1 ms instead of 2.3 ms per launch)
  const uint2* src = reinterpret_cast<const uint2*>(lo_area + 8192);
  uint2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = src[i * 128 + ctid];
  asm volatile("bar.sync 2, 128;" ::: "memory");
  uint4* dst = reinterpret_cast<uint4*>(lo_area);
  const __half2 bias = __floats2half2_rn(1152.f, 1152.f);
  auto widen = [&](uint32_t w, uint32_t& o0, uint32_t& o1) {
    const uint32_t p0 = __byte_perm(w, 0x64646464u, 0x4140), p1 = __byte_perm(w, 0x64646464u, 0x4342);
    const __half2 h0 = __hsub2(*reinterpret_cast<const __half2*>(&p0), bias), h1 = __hsub2(*reinterpret_cast<const __half2*>(&p1), bias);
    o0 = *reinterpret_cast<const uint32_t*>(&h0);
    o1 = *reinterpret_cast<const uint32_t*>(&h1);
  };
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 o;
    widen(a[i].x, o.x, o.y);
    widen(a[i].y, o.z, o.w);
    dst[i * 128 + ctid] = o;
  }
}

}  // namespace p16
}  // namespace odf
