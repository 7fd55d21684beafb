// Second half of a K^T (K v) sweep on the tensor cores: the fused tile (odf_gauss_tile.cu, SPILL16) leaves the K
// tiles of one row chunk in HBM as two planes, hi = rn16(K) (fp16) and lo = rni((K - hi) * 2^19) + 128 (ONE BYTE: the
// residual in fixed point, 2^-20 absolute on K <= 1 -- 3 B per kernel value instead of the 4 B of an fp16 lo plane,
// a quarter less HBM traffic for kernels that do nothing but stream the panel), tile-blocked as
//     P16[plane][column tile j][row block rb][group g of 8 centres][128 rows][8]      (32 KB / 16 KB per (j, rb))
// (the tile's epilogue owns one row per thread: with this order a warp-wide 16-byte store covers 512 contiguous bytes)
// and this kernel contracts them with the finished W = K v + w of the same rows,
//     out_partial[s][c][t] = sum_{r in row range s} K[r][c] * W[r][t],
// i.e. the K_blk^T w half of falkon `GaussianKernel.dmmv` (reached from InCoreFalkon.fit,
// src/modules/region-classifier/FALKONWrapper_with_centers_selection_incore.py:68).
//
// The product is 16 flop per panel byte — far below the tensor-pipe balance — so the kernel's job is to stream the
// panel from HBM once, at copy speed; the fp32-FMA version (odf_panel.cu) was bound by the FMA pipe at 2.6 TB/s.
// Here the centres are the M dimension of a tcgen05.mma kind::f16 (A = P^T, MN-major: a [8 rows x 8 centres] block of
// the layout above is exactly one un-swizzled 128-byte core matrix), the rows are its K dimension, and
// B = W16 [row][hi(s_t W) 0..31 | lo 32..63] (MN-major SWIZZLE_128B, one 128-byte row per panel row, s_t a power of two
// per column so that s_t max|W[:, t]| < 2^15):
//     acc1 [128 x 64] += P_hi^T . W16      (columns 0..31: hi.hi      32..63: hi.lo, scaled 2^11)
//     acc2 [128 x 32] += P_lo^T . W16[:, 0..31]   (lo.hi, scaled 2^19; an N = 32 MMA: lo.lo is never formed -- a quarter less
//                                                  tensor work, which under the 1 kW cap is ~50 MHz of SM clock and 3 % of the pass)
// fp32 in TMEM.  The tensor core adds with truncation, so an accumulation chain is cut every 512 rows: the epilogue
// warps drain the (double-buffered) accumulators into registers, combine the three terms and keep the running sums in
// fp32 round-to-nearest.  Row ranges write separate slabs that odf_finish_rows reduces in index order (deterministic).
//
// Persistent, warp-specialised (384 threads, 1 CTA / SM): warp 0 TMA producer (5 stages of 64 rows: 32 KB of loads each,
// 160 KB in flight per SM), warps 8-11 widen the stage's lo bytes to fp16 in place (widen_lo8), warp 1 MMA issuer, warp 2
// TMEM allocator, warps 4-7 epilogue (thread = TMEM lane = centre).
#include <cstdlib>
#include <cuda_fp16.h>
#include "odf_ptx.cuh"
#include "odf_internal.h"
#include "odf_panel16_common.cuh"

namespace odf {

namespace {

using namespace p16;

struct Panel16Params {
  int n_ct;               // column (centre) tiles of 128
  int n_rb;               // row blocks of 128 in the panel layout
  int n_stages;           // 64-row stages in the chunk
  int stages_per_item;
  int n_rsplit;           // row ranges (= partial slabs)
  int M, T_pad;
  int plane_rows;         // rows of the [.. x 1024] fp16 view (one per (j, rb, g)) between the hi and the lo plane
  int swap_lbo_sbo;       // bring-up switch (env ODF_P16_SWAP)
  const uint32_t* absmax; // 32 words: bits of max|W[:, t]| (fix the per-column power-of-two scales of W16)
  float* out;             // [n_rsplit][M][T_pad]
};

// HI_ONLY = 1 (experimental precision tier): only the hi plane is streamed (2 B per kernel value, 11 bits of
// K): the lo loads, the acc2 MMAs and the lo.hi term of the read-out are dropped.
template <int HI_ONLY>
__global__ void __launch_bounds__(QTHREADS, 1)
panel16_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmW,
               const Panel16Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + QNS * QSTAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + QBARS);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int B_FULL = 0, B_EMPTY = QNS, B_AFULL = 2 * QNS, B_AEMPTY = 2 * QNS + 2, B_CONV = 2 * QNS + 4;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmP);
    tma_prefetch_desc(&tmL);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < QNS; ++s) {
      mbar_init(BAR(B_FULL + s), 1);
      mbar_init(BAR(B_EMPTY + s), 1);
      mbar_init(BAR(B_CONV + s), 4);
    }
    mbar_init(BAR(B_AFULL + 0), 1);
    mbar_init(BAR(B_AFULL + 1), 1);
    mbar_init(BAR(B_AEMPTY + 0), 128);
    mbar_init(BAR(B_AEMPTY + 1), 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), QTM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = p.n_ct * p.n_rsplit;
  // item -> (row range s, column tile j), j fastest: CTAs running side by side share the W16 rows in L2

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int s = it / p.n_ct, j = it - s * p.n_ct;
        const int st0 = s * p.stages_per_item;
        const int st1 = min(st0 + p.stages_per_item, p.n_stages);
        for (int st = st0; st < st1; ++st) {
          mbar_wait(BAR(B_EMPTY + stage), phase ^ 1);
          const uint32_t full = BAR(B_FULL + stage);
          const uint32_t dst = smem_u32(smem) + stage * QSTAGE;
          // [total x 1024] fp16 view, one row per (plane, j, rb, g) holding [128 rows][8]; a box is 16 g x 32 rows
          const int y = (j * p.n_rb + (st >> 1)) * 16;
          const int x = (st & 1) * 512;
          mbar_arrive_expect_tx(full, HI_ONLY ? QSTAGE - 2 * QBOX : QSTAGE - QBOX);
          // the planes are read exactly once (evict first); W16 is shared by every column tile (evict last)
          tma_load_2d_hint(dst + 0 * QBOX, &tmP, full, x, y, kEvictFirst);
          tma_load_2d_hint(dst + 1 * QBOX, &tmP, full, x + 256, y, kEvictFirst);
          if (!HI_ONLY) {
            // lo bytes: two boxes of [16 centre groups][32 rows][8 B] = 4 KB, into the upper half of the fp16 lo area
            tma_load_2d_hint(dst + 3 * QBOX, &tmL, full, x, y, kEvictFirst);
            tma_load_2d_hint(dst + 3 * QBOX + QBOX / 2, &tmL, full, x + 256, y, kEvictFirst);
          }
          tma_load_2d_hint(dst + 4 * QBOX, &tmW, full, 0, st * QR, kEvictLast);
          if (++stage == QNS) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      const uint32_t sdesc0 = (smem_u32(smem) & 0x3FFFFu) >> 4;
      const uint64_t a_const = p.swap_lbo_sbo ? kSdescMnPlainHiSwapped : kSdescMnPlainHi;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t g = 0;                       // accumulation chains started so far
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int s = it / p.n_ct;
        const int st0 = s * p.stages_per_item;
        const int st1 = min(st0 + p.stages_per_item, p.n_stages);
        for (int st = st0; st < st1; ++st) {
          const int local = st - st0;
          const bool first = (local % QFLUSH) == 0;
          const uint32_t t_acc = tmem_base + (g & 1) * 128;
          if (first) mbar_wait(BAR(B_AEMPTY + (g & 1)), ((g >> 1) & 1) ^ 1);
          mbar_wait(BAR(B_FULL + stage), phase);
          if (!HI_ONLY) mbar_wait(BAR(B_CONV + stage), phase);     // the lo bytes have been widened to fp16
          tc_fence_after();
          const uint32_t sd = sdesc0 + stage * (QSTAGE >> 4);
#pragma unroll
          for (int kk = 0; kk < QR / 16; ++kk) {          // K = 16 rows per MMA: two core matrices of A, two 8-row groups of B
            const uint32_t a_off = (kk >> 1) * QBOX + (kk & 1) * 256;
            const uint64_t a_hi = a_const | static_cast<uint64_t>(sd + ((0 * QBOX + a_off) >> 4));
            const uint64_t a_lo = a_const | static_cast<uint64_t>(sd + ((2 * QBOX + a_off) >> 4));
            const uint64_t b_w = kSdescMnHi | static_cast<uint64_t>(sd + ((4 * QBOX + kk * 2048) >> 4));
            const uint32_t accum = (first && kk == 0) ? 0u : 1u;
            mma_f16_ss(t_acc, a_hi, b_w, kIdesc, accum);
            if (!HI_ONLY) mma_f16_ss(t_acc + 64, a_lo, b_w, kIdescN32, accum);
          }
          tc_commit(BAR(B_EMPTY + stage));
          if (((local + 1) % QFLUSH) == 0 || st == st1 - 1) {
            tc_commit(BAR(B_AFULL + (g & 1)));
            ++g;
          }
          if (++stage == QNS) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ======================= epilogue (warps 4-7) =======================
    const int q = warp & 3;
    const int row = q * 32 + lane;                        // centre inside the column tile (= TMEM lane)
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    uint32_t g = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int s = it / p.n_ct, j = it - s * p.n_ct;
      const int st0 = s * p.stages_per_item;
      const int st1 = min(st0 + p.stages_per_item, p.n_stages);
      const int n_chains = (st1 - st0 + QFLUSH - 1) / QFLUSH;
      float acc[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) acc[t] = 0.f;
      for (int c = 0; c < n_chains; ++c, ++g) {
        const uint32_t b = g & 1;
        mbar_wait_warp(BAR(B_AFULL + b), (g >> 1) & 1);
        tc_fence_after();
        const uint32_t t_acc = tmem_base + lane_off + b * 128;
        uint32_t r[32];
        float tmp[32];
        __syncwarp();
        if (!HI_ONLY) {
          tmem_ld32(t_acc + 64, r);                       // lo.hi  (x 2^-19)
          tc_wait_ld();
#pragma unroll
          for (int t = 0; t < 32; ++t) tmp[t] = __uint_as_float(r[t]) * kLoScale;
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) tmp[t] = 0.f;
        }
        tmem_ld32(t_acc + 32, r);                         // hi.lo  (x 2^-11)
        tc_wait_ld();
#pragma unroll
        for (int t = 0; t < 32; ++t) tmp[t] = fmaf(__uint_as_float(r[t]), 1.f / 2048.f, tmp[t]);
        tmem_ld32(t_acc, r);                              // hi.hi
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(BAR(B_AEMPTY + b));
#pragma unroll
        for (int t = 0; t < 32; ++t) acc[t] += __uint_as_float(r[t]) + tmp[t];
      }
      const int m = j * 128 + row;
      if (m < p.M) {
        float* orow = p.out + (static_cast<int64_t>(s) * p.M + m) * p.T_pad;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          if (4 * v < p.T_pad) {
            const uint4 mx = __ldg(reinterpret_cast<const uint4*>(p.absmax) + v);     // per-column scales of W16
            float4 t4;
            t4.x = acc[4 * v + 0] * w16_scale_from_bits(mx.x, true); t4.y = acc[4 * v + 1] * w16_scale_from_bits(mx.y, true);
            t4.z = acc[4 * v + 2] * w16_scale_from_bits(mx.z, true); t4.w = acc[4 * v + 3] * w16_scale_from_bits(mx.w, true);
            *reinterpret_cast<float4*>(orow + 4 * v) = t4;
          }
        }
      }
    }
  } else if (warp >= 8) {
    // ======================= converters (warps 8-11): lo bytes -> fp16, in place =======================
    if (!HI_ONLY) {
      const int ctid = static_cast<int>(threadIdx.x) - 256;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int s = it / p.n_ct;
        const int st0 = s * p.stages_per_item;
        const int st1 = min(st0 + p.stages_per_item, p.n_stages);
        for (int st = st0; st < st1; ++st) {
          mbar_wait_warp(BAR(B_FULL + stage), phase);
          widen_lo8(smem + stage * QSTAGE + 2 * QBOX, ctid);
          fence_proxy_async();                            // generic-proxy writes before the tensor core's reads
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_CONV + stage));
          if (++stage == QNS) { stage = 0; phase ^= 1; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, QTM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// K . V from the SAME panel (the first half of a sweep when the panels are resident in HBM): the centres are the K
// dimension now and the rows the M dimension of the MMA.  A [8 rows x 8 centres] block of the panel layout is also a
// K-major un-swizzled core matrix (8 rows of 16 bytes), and the 16 centre groups of one (column tile, row block) are
// 32 KB of contiguous memory per plane, so a stage of 64 centres is two plain 16 KB bulk copies (hi, lo) that land in
// shared memory exactly as the tensor core wants them: next 8 rows +128 B (SBO), next group of 8 centres +2048 B (LBO).
//     out_partial[s][r][t] = sum_{c in column range s} K[r][c] * V[c][t]
// B = V16 [centre][hi(s_t V) 0..31 | lo 32..63], the same MN-major SWIZZLE_128B operand as W16 above.  Same pipeline,
// accumulators, chain length (512 centres) and epilogue (thread = TMEM lane = row) as panel16_kernel.
struct Panel16VParams {
  int n_ct;               // column (centre) tiles of 128
  int n_rb;               // row blocks of 128
  int n_stages;           // 64-centre stages per row block (2 per column tile)
  int stages_per_item;
  int n_csplit;           // column ranges (= partial slabs)
  int n_rows, T_pad;
  int64_t plane_elems;    // fp16 elements between the hi and the lo plane
  int swap_lbo_sbo;       // bring-up switch (env ODF_P16V_SWAP)
  const __half* P;        // the panel
  const uint32_t* absmax; // 32 words: bits of max|V[:, t]|
  float* out;             // [n_csplit][n_rows][T_pad]
};

template <int HI_ONLY>
__global__ void __launch_bounds__(QTHREADS, 1)
panel16_mmv_kernel(const __grid_constant__ CUtensorMap tmV, const Panel16VParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + QNS * VSTAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + QBARS);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int B_FULL = 0, B_EMPTY = QNS, B_AFULL = 2 * QNS, B_AEMPTY = 2 * QNS + 2, B_CONV = 2 * QNS + 4;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmV);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < QNS; ++s) {
      mbar_init(BAR(B_FULL + s), 1);
      mbar_init(BAR(B_EMPTY + s), 1);
      mbar_init(BAR(B_CONV + s), 4);
    }
    mbar_init(BAR(B_AFULL + 0), 1);
    mbar_init(BAR(B_AFULL + 1), 1);
    mbar_init(BAR(B_AEMPTY + 0), 128);
    mbar_init(BAR(B_AEMPTY + 1), 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), QTM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = p.n_rb * p.n_csplit;
  // item -> (column range s, row block rb), rb fastest: CTAs running side by side share the V16 rows in L2

  if (warp == 0) {
    // ======================= producer =======================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int s = it / p.n_rb, rb = it - s * p.n_rb;
        const int st0 = s * p.stages_per_item;
        const int st1 = min(st0 + p.stages_per_item, p.n_stages);
        for (int st = st0; st < st1; ++st) {
          mbar_wait(BAR(B_EMPTY + stage), phase ^ 1);
          const uint32_t full = BAR(B_FULL + stage);
          const uint32_t dst = smem_u32(smem) + stage * VSTAGE;
          // 8 centre groups of column tile j = st / 2, half st % 2: 16 KB of contiguous panel per plane
          const __half* src = p.P + ((static_cast<int64_t>(st >> 1) * p.n_rb + rb) * 16 + (st & 1) * 8) * 1024;
          mbar_arrive_expect_tx(full, HI_ONLY ? VSTAGE - VA : VSTAGE - VA / 2);
          bulk_g2s_hint(dst, src, VA, full, kEvictFirst);
          // lo bytes of the same 8 centre groups: 8 KB, into the upper half of the fp16 lo area
          if (!HI_ONLY) bulk_g2s_hint(dst + VA + VA / 2, reinterpret_cast<const uint8_t*>(p.P) + 2 * p.plane_elems + (src - p.P), VA / 2, full, kEvictFirst);
          tma_load_2d_hint(dst + 2 * VA, &tmV, full, 0, st * QR, kEvictLast);
          if (++stage == QNS) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      const uint32_t sdesc0 = (smem_u32(smem) & 0x3FFFFu) >> 4;
      const uint64_t a_const = p.swap_lbo_sbo ? kSdescKPlainHiSwapped : kSdescKPlainHi;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t g = 0;                       // accumulation chains started so far
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int s = it / p.n_rb;
        const int st0 = s * p.stages_per_item;
        const int st1 = min(st0 + p.stages_per_item, p.n_stages);
        for (int st = st0; st < st1; ++st) {
          const int local = st - st0;
          const bool first = (local % QFLUSH) == 0;
          const uint32_t t_acc = tmem_base + (g & 1) * 128;
          if (first) mbar_wait(BAR(B_AEMPTY + (g & 1)), ((g >> 1) & 1) ^ 1);
          mbar_wait(BAR(B_FULL + stage), phase);
          if (!HI_ONLY) mbar_wait(BAR(B_CONV + stage), phase);     // the lo bytes have been widened to fp16
          tc_fence_after();
          const uint32_t sd = sdesc0 + stage * (VSTAGE >> 4);
#pragma unroll
          for (int kk = 0; kk < QR / 16; ++kk) {          // K = 16 centres per MMA: two centre groups of A, two 8-row groups of B
            const uint64_t a_hi = a_const | static_cast<uint64_t>(sd + ((kk * 4096) >> 4));
            const uint64_t a_lo = a_const | static_cast<uint64_t>(sd + ((VA + kk * 4096) >> 4));
            const uint64_t b_v = kSdescMnHi | static_cast<uint64_t>(sd + ((2 * VA + kk * 2048) >> 4));
            const uint32_t accum = (first && kk == 0) ? 0u : 1u;
            mma_f16_ss(t_acc, a_hi, b_v, kIdescV, accum);
            if (!HI_ONLY) mma_f16_ss(t_acc + 64, a_lo, b_v, kIdescVN32, accum);
          }
          tc_commit(BAR(B_EMPTY + stage));
          if (((local + 1) % QFLUSH) == 0 || st == st1 - 1) {
            tc_commit(BAR(B_AFULL + (g & 1)));
            ++g;
          }
          if (++stage == QNS) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ======================= epilogue (warps 4-7) =======================
    const int q = warp & 3;
    const int row = q * 32 + lane;                        // row inside the row block (= TMEM lane)
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    uint32_t g = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int s = it / p.n_rb, rb = it - s * p.n_rb;
      const int st0 = s * p.stages_per_item;
      const int st1 = min(st0 + p.stages_per_item, p.n_stages);
      const int n_chains = (st1 - st0 + QFLUSH - 1) / QFLUSH;
      float acc[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) acc[t] = 0.f;
      for (int c = 0; c < n_chains; ++c, ++g) {
        const uint32_t b = g & 1;
        mbar_wait_warp(BAR(B_AFULL + b), (g >> 1) & 1);
        tc_fence_after();
        const uint32_t t_acc = tmem_base + lane_off + b * 128;
        uint32_t r[32];
        float tmp[32];
        __syncwarp();
        if (!HI_ONLY) {
          tmem_ld32(t_acc + 64, r);                       // lo.hi  (x 2^-19)
          tc_wait_ld();
#pragma unroll
          for (int t = 0; t < 32; ++t) tmp[t] = __uint_as_float(r[t]) * kLoScale;
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) tmp[t] = 0.f;
        }
        tmem_ld32(t_acc + 32, r);                         // hi.lo  (x 2^-11)
        tc_wait_ld();
#pragma unroll
        for (int t = 0; t < 32; ++t) tmp[t] = fmaf(__uint_as_float(r[t]), 1.f / 2048.f, tmp[t]);
        tmem_ld32(t_acc, r);                              // hi.hi
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(BAR(B_AEMPTY + b));
#pragma unroll
        for (int t = 0; t < 32; ++t) acc[t] += __uint_as_float(r[t]) + tmp[t];
      }
      const int m = rb * 128 + row;
      if (m < p.n_rows) {
        float* orow = p.out + (static_cast<int64_t>(s) * p.n_rows + m) * p.T_pad;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          if (4 * v < p.T_pad) {
            const uint4 mx = __ldg(reinterpret_cast<const uint4*>(p.absmax) + v);     // per-column scales of V16
            float4 t4;
            t4.x = acc[4 * v + 0] * w16_scale_from_bits(mx.x, true); t4.y = acc[4 * v + 1] * w16_scale_from_bits(mx.y, true);
            t4.z = acc[4 * v + 2] * w16_scale_from_bits(mx.z, true); t4.w = acc[4 * v + 3] * w16_scale_from_bits(mx.w, true);
            *reinterpret_cast<float4*>(orow + 4 * v) = t4;
          }
        }
      }
    }
  } else if (warp >= 8) {
    // ======================= converters (warps 8-11): lo bytes -> fp16, in place =======================
    if (!HI_ONLY) {
      const int ctid = static_cast<int>(threadIdx.x) - 256;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int s = it / p.n_rb;
        const int st0 = s * p.stages_per_item;
        const int st1 = min(st0 + p.stages_per_item, p.n_stages);
        for (int st = st0; st < st1; ++st) {
          mbar_wait_warp(BAR(B_FULL + stage), phase);
          widen_lo8(smem + stage * VSTAGE + VA, ctid);
          fence_proxy_async();                            // generic-proxy writes before the tensor core's reads
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_CONV + stage));
          if (++stage == QNS) { stage = 0; phase ^= 1; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, QTM_COLS);
}

int q_num_sms() { return device_sm_count(); }

}  // namespace

// bytes of the panel of one chunk: padded rows x padded centres x (2 B hi plane + 1 B lo plane)
size_t panel16_bytes(int64_t n_rows, int64_t M) {
  return static_cast<size_t>(round_up(n_rows, 128)) * static_cast<size_t>(round_up(M, 128)) * 3;
}

// Row ranges per column tile: chosen so that the slowest CTA of the persistent grid streams as few stages as possible
// (ceil(items / SMs) * stages_per_item), ranges of at least 8 stages, at most 32 slabs.
int panel16_splits(int64_t n_rows, int64_t M) {
  const int64_t n_ct = (M + 127) / 128, n_st = (n_rows + QR - 1) / QR;
  const int sms = 148;                                   // fixed so that the slab count does not depend on the device
  int64_t best = 1, best_cost = INT64_MAX;
  for (int64_t s = 1; s <= 32; ++s) {
    const int64_t spi = (n_st + s - 1) / s;
    if (s > 1 && spi < 8) break;
    const int64_t s_eff = (n_st + spi - 1) / spi;
    const int64_t items = n_ct * s_eff;
    const int64_t cost = (items + sms - 1) / sms * spi;
    if (cost < best_cost) { best_cost = cost; best = s_eff; }
  }
  return static_cast<int>(best);
}

int launch_panel16_tmm(const void* P16, int64_t n_rows, int64_t M, const void* W16, const uint32_t* absmax, int T_pad,
                       int n_splits, float* out_partial, cudaStream_t st, int hi_only) {
  if (n_rows <= 0 || M <= 0 || (T_pad != 16 && T_pad != 32) || (reinterpret_cast<uintptr_t>(P16) & 127) != 0 ||
      (reinterpret_cast<uintptr_t>(W16) & 127) != 0 || (reinterpret_cast<uintptr_t>(out_partial) & 15) != 0 || !absmax || (reinterpret_cast<uintptr_t>(absmax) & 15) != 0)
    return set_error(ODF_ERR_ARG, "panel16_tmm: bad shape or alignment (P16, W16 128-byte aligned; T_pad 16 or 32)");
  if (n_splits != panel16_splits(n_rows, M)) return set_error(ODF_ERR_ARG, "panel16_tmm: n_splits must come from odf_panel16_splits");
  static DeviceOnce attr_once;
  bool& attr_set = attr_once.here();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(panel16_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, QSMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(panel16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, QSMEM);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(panel16_kernel)");
    attr_set = true;
  }
  Panel16Params p;
  p.n_ct = static_cast<int>((M + 127) / 128);
  p.n_rb = static_cast<int>((n_rows + 127) / 128);
  p.n_stages = static_cast<int>((n_rows + QR - 1) / QR);
  p.stages_per_item = (p.n_stages + n_splits - 1) / n_splits;
  p.n_rsplit = n_splits;
  p.M = static_cast<int>(M);
  p.T_pad = T_pad;
  const int64_t plane_rows = static_cast<int64_t>(p.n_ct) * p.n_rb * 16;
  if (plane_rows > 0x7fffffffll) return set_error(ODF_ERR_ARG, "panel16_tmm: chunk too large for 32-bit TMA coordinates");
  p.plane_rows = static_cast<int>(plane_rows);
  {
    const char* e = getenv("ODF_P16_SWAP");
    p.swap_lbo_sbo = e ? atoi(e) : 0;
  }
  p.absmax = absmax;
  p.out = out_partial;
  CUtensorMap tmP, tmL, tmW;
  int rc;
  if ((rc = make_map_plain_f16(&tmP, P16, plane_rows, 1024, 1024, 16, 256))) return rc;
  if ((rc = make_map_plain_u8(&tmL, static_cast<const uint8_t*>(P16) + plane_rows * 2048, plane_rows, 1024, 1024, 16, 256))) return rc;
  if ((rc = make_map_sw128(&tmW, W16, static_cast<int64_t>(p.n_rb) * 128, 64, 64, QR, 2))) return rc;
  const int n_items = p.n_ct * p.n_rsplit;
  const int sms = q_num_sms();
  const int grid = n_items < sms ? n_items : sms;
  if (hi_only) panel16_kernel<1><<<grid, QTHREADS, QSMEM, st>>>(tmP, tmL, tmW, p);
  else panel16_kernel<0><<<grid, QTHREADS, QSMEM, st>>>(tmP, tmL, tmW, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "panel16_kernel launch");
  return ODF_OK;
}

// Column ranges per row block for panel16_mmv: the same cost model with the roles of rows and centres swapped
// (items = row blocks, stages = 64-centre halves of the column tiles).
int panel16_mmv_splits(int64_t n_rows, int64_t M) { return panel16_splits(round_up(M, 128), n_rows); }

int launch_panel16_mmv(const void* P16, int64_t n_rows, int64_t M, const void* V16, const uint32_t* absmax, int T_pad,
                       int n_splits, float* out_partial, cudaStream_t st, int hi_only) {
  if (n_rows <= 0 || M <= 0 || (T_pad != 16 && T_pad != 32) || (reinterpret_cast<uintptr_t>(P16) & 127) != 0 ||
      (reinterpret_cast<uintptr_t>(V16) & 127) != 0 || (reinterpret_cast<uintptr_t>(out_partial) & 15) != 0 || !absmax || (reinterpret_cast<uintptr_t>(absmax) & 15) != 0)
    return set_error(ODF_ERR_ARG, "panel16_mmv: bad shape or alignment (P16, V16 128-byte aligned; T_pad 16 or 32)");
  if (n_splits != panel16_mmv_splits(n_rows, M)) return set_error(ODF_ERR_ARG, "panel16_mmv: n_splits must come from odf_panel16_mmv_splits");
  static DeviceOnce attr_once;
  bool& attr_set = attr_once.here();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(panel16_mmv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, QSMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(panel16_mmv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, QSMEM);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(panel16_mmv_kernel)");
    attr_set = true;
  }
  Panel16VParams p;
  p.n_ct = static_cast<int>((M + 127) / 128);
  p.n_rb = static_cast<int>((n_rows + 127) / 128);
  p.n_stages = 2 * p.n_ct;
  p.stages_per_item = (p.n_stages + n_splits - 1) / n_splits;
  p.n_csplit = n_splits;
  p.n_rows = static_cast<int>(n_rows);
  p.T_pad = T_pad;
  p.plane_elems = static_cast<int64_t>(p.n_ct) * p.n_rb * 16384;
  {
    const char* e = getenv("ODF_P16V_SWAP");
    p.swap_lbo_sbo = e ? atoi(e) : 0;
  }
  p.P = static_cast<const __half*>(P16);
  p.absmax = absmax;
  p.out = out_partial;
  CUtensorMap tmV;
  int rc;
  if ((rc = make_map_sw128(&tmV, V16, static_cast<int64_t>(p.n_ct) * 128, 64, 64, QR, 2))) return rc;
  const int n_items = p.n_rb * p.n_csplit;
  const int sms = q_num_sms();
  const int grid = n_items < sms ? n_items : sms;
  if (hi_only) panel16_mmv_kernel<1><<<grid, QTHREADS, QSMEM, st>>>(tmV, p);
  else panel16_mmv_kernel<0><<<grid, QTHREADS, QSMEM, st>>>(tmV, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "panel16_mmv_kernel launch");
  return ODF_OK;
}

}  // namespace odf
