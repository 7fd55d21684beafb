// One pass over a RESIDENT fp16-plane panel per K^T (K v) sweep.
//
// odf_panel16.cu forms K v and K^T w as two kernels, each streaming the whole panel of a row chunk from HBM (2 x 4 B per
// kernel value and sweep; 21 of the 23 sweeps of a fit, reached from InCoreFalkon.fit -> falkon `GaussianKernel.dmmv`,
// src/modules/region-classifier/FALKONWrapper_with_centers_selection_incore.py:68).  The two contractions cannot share
// one trip through shared memory -- K^T w needs w = K v of a row COMPLETE over all centres before the row is touched
// again, and a row of the panel is 40 KB -- but they can share one trip through the 126 MB L2: this kernel walks the
// chunk in groups of RG = 4 row blocks (512 rows, 20.7 MB of panel at M = 10 k) and runs, in ONE persistent grid,
//     A items  (group g, row block rb, centre range cs):   partial[cs][rows of rb] = K[rb, cs] . V16        (HBM -> L2 -> SM)
//     finish   (last A item of a row block to arrive):     w = sum_cs partial (fixed order), per-row-block scales,
//                                                          W16[rb] = fp16 hi | lo split of w
//     C items  (group g, column tile j):                   out[g][j] = K[g, j]^T . W16[g]                   (L2 -> SM)
// with the C items of group g scheduled LAG groups after its A items: by then the group's panel lines are still in L2
// (3 groups = 62 MB live at LAG = 2) and the row blocks' W16 are ready, so the second touch costs no HBM traffic.
// Both item types stream 8 stages of 40 KB through the same 5-stage ring, accumulate in the same double-buffered TMEM
// accumulators and use the MMA / epilogue code of the two-pass kernels; items are dealt round-robin (item i -> CTA
// i mod grid; all items are the same size, so the CTAs advance in lock step and a C item finds its flag set).
//
// Hand-offs through global memory (all CTAs are co-resident: grid <= SMs, 1 CTA / SM):
//   * A partials: ring of RING groups [slot][cs][512 rows][32] fp32 (stays in L2); the epilogue warps fence, count the
//     row block's arrivals with an atomic and the LAST arriver reduces the n_cs slabs in index order (deterministic).
//   * W16 [rows][hi 0..31 | lo 32..63] and wscale[rb][32] (inverse of the power-of-two scale of the row block: a chain
//     of the C items' accumulation is one row block = 2 stages); then flag[g] += 1, release; C items acquire flag[g] = 4.
//   * C partials: out[g][M][T_pad], one slab per group, reduced in index order by odf_finish_rows.
#include <cstdlib>
#include <cuda_fp16.h>
#include "odf_ptx.cuh"
#include "odf_internal.h"
#include "odf_panel16_common.cuh"

namespace odf {

namespace {

using namespace p16;

constexpr int RG = 4;                     // row blocks (of 128 rows) per group
constexpr int SW_STAGES = 8;              // stages per item (A: 8 x 64 centres of one row block; C: 8 x 64 rows of one column tile)
constexpr int SW_RING = 8;                // groups of A partials kept (slot = g mod SW_RING)
constexpr int SW_SMEM = QSMEM + 1024;     // + hand-off words of the epilogue warps

struct SweepParams {
  int n_ct, n_rb, n_groups;
  int n_cs;               // A items per row block
  int n_cstages;          // 64-centre stages per row block (2 per column tile)
  int lag;                // C items of group g are dealt with the A items of group g + lag
  int ipb;                // item slots per schedule block: RG * n_cs + n_ct
  int n_rows, M, T_pad;
  int plane_rows;         // TMA view rows between the hi and the lo plane
  int policy_a;           // 0: no hint, 1: evict_last, 2: evict_first for the first touch of the planes
  int policy_c;           // the same for the second touch
  int dbg_ncs;            // > 0: the finishing step reads only this many partial slabs (TIMING EXPERIMENTS ONLY: wrong results)
  int64_t plane_elems;    // fp16 elements between the hi and the lo plane
  const __half* P;
  const uint32_t* absmax_v;
  float* apart;           // [SW_RING][n_cs][RG * 128][32]
  __half* W16;            // [n_rb * 128][64]
  float* wscale;          // [n_rb][32]
  unsigned int* cnt;      // [n_rb]      A items of the row block that have arrived
  unsigned int* flag;     // [n_groups]  row blocks of the group whose W16 is ready
  float* out;             // [n_groups][M][T_pad]
};

struct Item {
  int type;               // 0: empty slot, 1: A, 2: C
  int g, rb, x;           // A: x = centre range cs;  C: x = column tile j, rb = first row block of the group
  int st0, st1;           // A: 64-centre stages [st0, st1) of the row block;  C: 64-row stages [st0, st1) of the chunk
};

__device__ __forceinline__ Item decode_item(const SweepParams& p, int it) {
  Item I;
  I.type = 0; I.g = 0; I.rb = 0; I.x = 0; I.st0 = 0; I.st1 = 0;
  const int b = it / p.ipb, pos = it - b * p.ipb;
  const int na = RG * p.n_cs;
  if (pos < na) {
    const int rbi = pos / p.n_cs, cs = pos - rbi * p.n_cs;
    const int rb = b * RG + rbi;
    if (b < p.n_groups && rb < p.n_rb) {
      I.type = 1; I.g = b; I.rb = rb; I.x = cs;
      I.st0 = cs * SW_STAGES;
      I.st1 = min(I.st0 + SW_STAGES, p.n_cstages);
    }
  } else {
    const int g = b - p.lag;
    if (g >= 0 && g < p.n_groups) {
      const int rbs = min(RG, p.n_rb - g * RG);
      I.type = 2; I.g = g; I.rb = g * RG; I.x = pos - na;
      I.st0 = g * RG * 2;
      I.st1 = I.st0 + 2 * rbs;
    }
  }
  return I;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// generic-proxy writes of other threads (made visible by their release / our acquire) before async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// spin until *f >= need; a protocol bug traps instead of hanging the device
__device__ __forceinline__ void wait_flag(const unsigned int* f, unsigned int need) {
  if (ld_acquire_u32(f) >= need) return;
  const long long t0 = clock64();
  while (ld_acquire_u32(f) < need) {
    __nanosleep(100);
    if ((clock64() - t0) > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ uint64_t policy_of(int k) { return k == 1 ? kEvictLast : kEvictFirst; }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(256, 1)
panel16_sweep_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmW, const SweepParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + QNS * QSTAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + QBARS);
  uint32_t* hand = tmem_slot + 4;           // [0]: last-arriver flag; [4 .. 4+128): per-warp column maxima; [132 .. 164): maxima
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int B_FULL = 0, B_EMPTY = QNS, B_AFULL = 2 * QNS, B_AEMPTY = 2 * QNS + 2;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmP);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < QNS; ++s) {
      mbar_init(BAR(B_FULL + s), 1);
      mbar_init(BAR(B_EMPTY + s), 1);
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

  const int n_items = (p.n_groups + p.lag) * p.ipb;

  if (warp == 0) {
    // ======================= producer =======================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      unsigned int flag_c = 0u;             // flag of THIS item's group as read while the previous item was being issued
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item I = decode_item(p, it);
        if (I.type == 0) continue;
        // look ahead: the flag of the next C item is read now and used after this item's stages have been issued
        // (an L2 round trip that would otherwise sit between two items with nothing being issued)
        int it_n = it + gridDim.x;
        Item In = decode_item(p, it_n < n_items ? it_n : 0);
        while (it_n < n_items && In.type == 0) { it_n += gridDim.x; In = decode_item(p, it_n < n_items ? it_n : 0); }
        unsigned int flag_n = 0u;
        const bool peek = it_n < n_items && In.type == 2;
        if (peek) flag_n = ld_acquire_u32(p.flag + In.g);
        if (I.type == 1) {
          for (int st = I.st0; st < I.st1; ++st) {
            mbar_wait(BAR(B_EMPTY + stage), phase ^ 1);
            const uint32_t full = BAR(B_FULL + stage);
            const uint32_t dst = smem_u32(smem) + stage * VSTAGE;
            // 8 centre groups of column tile st / 2, half st % 2, of row block rb: 16 KB of contiguous panel per plane
            const __half* src = p.P + ((static_cast<int64_t>(st >> 1) * p.n_rb + I.rb) * 16 + (st & 1) * 8) * 1024;
            mbar_arrive_expect_tx(full, VSTAGE);
            if (p.policy_a) {
              bulk_g2s_hint(dst, src, VA, full, policy_of(p.policy_a));
              bulk_g2s_hint(dst + VA, src + p.plane_elems, VA, full, policy_of(p.policy_a));
            } else {
              bulk_g2s(dst, src, VA, full);
              bulk_g2s(dst + VA, src + p.plane_elems, VA, full);
            }
            tma_load_2d_hint(dst + 2 * VA, &tmV, full, 0, st * QR, kEvictLast);
            if (++stage == QNS) { stage = 0; phase ^= 1; }
          }
        } else {
          // the W16 rows of the group are written by other CTAs: acquire, then order the TMA reads behind it
          if (flag_c < static_cast<unsigned int>((I.st1 - I.st0) >> 1)) wait_flag(p.flag + I.g, static_cast<unsigned int>((I.st1 - I.st0) >> 1));
          fence_proxy_async_all();
          for (int st = I.st0; st < I.st1; ++st) {
            mbar_wait(BAR(B_EMPTY + stage), phase ^ 1);
            const uint32_t full = BAR(B_FULL + stage);
            const uint32_t dst = smem_u32(smem) + stage * QSTAGE;
            const int y = (I.x * p.n_rb + (st >> 1)) * 16;
            const int x = (st & 1) * 512;
            mbar_arrive_expect_tx(full, QSTAGE);
            if (p.policy_c) {
              const uint64_t pol = policy_of(p.policy_c);
              tma_load_2d_hint(dst + 0 * QBOX, &tmP, full, x, y, pol);
              tma_load_2d_hint(dst + 1 * QBOX, &tmP, full, x + 256, y, pol);
              tma_load_2d_hint(dst + 2 * QBOX, &tmP, full, x, p.plane_rows + y, pol);
              tma_load_2d_hint(dst + 3 * QBOX, &tmP, full, x + 256, p.plane_rows + y, pol);
            } else {
              tma_load_2d(dst + 0 * QBOX, &tmP, full, x, y);
              tma_load_2d(dst + 1 * QBOX, &tmP, full, x + 256, y);
              tma_load_2d(dst + 2 * QBOX, &tmP, full, x, p.plane_rows + y);
              tma_load_2d(dst + 3 * QBOX, &tmP, full, x + 256, p.plane_rows + y);
            }
            tma_load_2d_hint(dst + 4 * QBOX, &tmW, full, 0, st * QR, kEvictLast);
            if (++stage == QNS) { stage = 0; phase ^= 1; }
          }
        }
        flag_c = peek ? flag_n : 0u;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      const uint32_t sdesc0 = (smem_u32(smem) & 0x3FFFFu) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t g = 0;                       // accumulation chains started so far
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item I = decode_item(p, it);
        if (I.type == 0) continue;
        const int chain = (I.type == 1) ? SW_STAGES : 2;      // A: the whole item; C: one row block (its own W16 scale)
        for (int st = I.st0; st < I.st1; ++st) {
          const int local = st - I.st0;
          const bool first = (local % chain) == 0;
          const uint32_t t_acc = tmem_base + (g & 1) * 128;
          if (first) mbar_wait(BAR(B_AEMPTY + (g & 1)), ((g >> 1) & 1) ^ 1);
          mbar_wait(BAR(B_FULL + stage), phase);
          tc_fence_after();
          const uint32_t sd = sdesc0 + stage * (QSTAGE >> 4);
          if (I.type == 1) {
#pragma unroll
            for (int kk = 0; kk < QR / 16; ++kk) {          // K = 16 centres per MMA
              const uint64_t a_hi = kSdescKPlainHi | static_cast<uint64_t>(sd + ((kk * 4096) >> 4));
              const uint64_t a_lo = kSdescKPlainHi | static_cast<uint64_t>(sd + ((VA + kk * 4096) >> 4));
              const uint64_t b_v = kSdescMnHi | static_cast<uint64_t>(sd + ((2 * VA + kk * 2048) >> 4));
              const uint32_t accum = (first && kk == 0) ? 0u : 1u;
              mma_f16_ss(t_acc, a_hi, b_v, kIdescV, accum);
              mma_f16_ss(t_acc + 64, a_lo, b_v, kIdescV, accum);
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < QR / 16; ++kk) {          // K = 16 rows per MMA
              const uint32_t a_off = (kk >> 1) * QBOX + (kk & 1) * 256;
              const uint64_t a_hi = kSdescMnPlainHi | static_cast<uint64_t>(sd + ((0 * QBOX + a_off) >> 4));
              const uint64_t a_lo = kSdescMnPlainHi | static_cast<uint64_t>(sd + ((2 * QBOX + a_off) >> 4));
              const uint64_t b_w = kSdescMnHi | static_cast<uint64_t>(sd + ((4 * QBOX + kk * 2048) >> 4));
              const uint32_t accum = (first && kk == 0) ? 0u : 1u;
              mma_f16_ss(t_acc, a_hi, b_w, kIdesc, accum);
              mma_f16_ss(t_acc + 64, a_lo, b_w, kIdesc, accum);
            }
          }
          tc_commit(BAR(B_EMPTY + stage));
          if (((local + 1) % chain) == 0 || st == I.st1 - 1) {
            tc_commit(BAR(B_AFULL + (g & 1)));
            ++g;
          }
          if (++stage == QNS) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ======================= epilogue (+ the finishing of a row block by its last A item) =======================
    const int q = warp & 3;
    const int row = q * 32 + lane;                        // TMEM lane: row of the row block (A) / centre of the tile (C)
    const int etid = row;                                 // thread index inside the epilogue group
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    uint32_t g = 0;
    auto drain = [&](float (&val)[32]) {                  // val = hi.hi + hi.lo 2^-11 + lo.hi 2^-12 of the next chain
      const uint32_t b = g & 1;
      mbar_wait_warp(BAR(B_AFULL + b), (g >> 1) & 1);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + lane_off + b * 128;
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(t_acc + 64, r);                           // lo.hi  (x 2^-12)
      tc_wait_ld();
#pragma unroll
      for (int t = 0; t < 32; ++t) val[t] = __uint_as_float(r[t]) * (1.f / 4096.f);
      tmem_ld32(t_acc + 32, r);                           // hi.lo  (x 2^-11)
      tc_wait_ld();
#pragma unroll
      for (int t = 0; t < 32; ++t) val[t] = fmaf(__uint_as_float(r[t]), 1.f / 2048.f, val[t]);
      tmem_ld32(t_acc, r);                                // hi.hi
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(BAR(B_AEMPTY + b));
#pragma unroll
      for (int t = 0; t < 32; ++t) val[t] += __uint_as_float(r[t]);
      ++g;
    };
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const Item I = decode_item(p, it);
      if (I.type == 0) continue;
      if (I.type == 1) {
        // ---- A item: partial sums of K v over the item's centres, one row per thread ----
        float val[32];
        drain(val);
        const int slot = I.g % SW_RING;
        const int rbi = I.rb - I.g * RG;
        if (I.g >= SW_RING) {
          // the slot's previous user (group g - RING) must have been reduced: true long ago in a healthy schedule
          if (lane == 0) wait_flag(p.flag + (I.g - SW_RING), static_cast<unsigned int>(RG));
          __syncwarp();
        }
        float* prow = p.apart + ((static_cast<int64_t>(slot) * p.n_cs + I.x) * (RG * 128) + rbi * 128 + row) * 32;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const uint4 mx = __ldg(reinterpret_cast<const uint4*>(p.absmax_v) + v);       // per-column scales of V16
          float4 t4;
          t4.x = val[4 * v + 0] * w16_scale_from_bits(mx.x, true); t4.y = val[4 * v + 1] * w16_scale_from_bits(mx.y, true);
          t4.z = val[4 * v + 2] * w16_scale_from_bits(mx.z, true); t4.w = val[4 * v + 3] * w16_scale_from_bits(mx.w, true);
          __stcg(reinterpret_cast<float4*>(prow) + v, t4);
        }
        __threadfence();
        epi_bar();
        if (etid == 0) {
          __threadfence();                                  // cumulative over the partial rows the barrier made visible to this thread
          const unsigned int old = atomicAdd(p.cnt + I.rb, 1u);
          hand[0] = (old == static_cast<unsigned int>(p.n_cs - 1)) ? 1u : 0u;
        }
        epi_bar();
        const bool last = hand[0] != 0u;
        if (last) {
          // ---- finish the row block: w = sum of the n_cs partials in index order, scales, fp16 split ----
          __threadfence();
          float w[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) w[t] = 0.f;
          const float* src = p.apart + (static_cast<int64_t>(slot) * p.n_cs * (RG * 128) + rbi * 128 + row) * 32;
          const int ncs_read = p.dbg_ncs > 0 ? min(p.dbg_ncs, p.n_cs) : p.n_cs;
          for (int cs = 0; cs < ncs_read; ++cs) {
            const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<int64_t>(cs) * (RG * 128) * 32);
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              const float4 t4 = __ldcg(s4 + v);
              w[4 * v + 0] += t4.x; w[4 * v + 1] += t4.y; w[4 * v + 2] += t4.z; w[4 * v + 3] += t4.w;
            }
          }
          const bool valid = (I.rb * 128 + row) < p.n_rows;
          if (!valid) {
#pragma unroll
            for (int t = 0; t < 32; ++t) w[t] = 0.f;
          }
          // per-column max |w| over the 128 rows (bit patterns of non-negative floats order as integers)
          uint32_t mine = 0u;
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(fabsf(w[t])));
            if (lane == t) mine = m;
          }
          hand[4 + q * 32 + lane] = mine;
          epi_bar();
          if (etid < 32) {
            const uint32_t m = max(max(hand[4 + etid], hand[4 + 32 + etid]), max(hand[4 + 64 + etid], hand[4 + 96 + etid]));
            hand[132 + etid] = m;
            p.wscale[static_cast<int64_t>(I.rb) * 32 + etid] = w16_scale_from_bits(m, true);
          }
          epi_bar();
          __align__(16) __half hi[32], lo[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const float sv = w[t] * w16_scale_from_bits(hand[132 + t], false);
            const __half h = __float2half_rn(sv);
            hi[t] = h;
            lo[t] = __float2half_rn((sv - __half2float(h)) * 2048.f);
          }
          uint4* wrow = reinterpret_cast<uint4*>(p.W16 + (static_cast<int64_t>(I.rb) * 128 + row) * 64);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            wrow[v] = reinterpret_cast<const uint4*>(hi)[v];
            wrow[4 + v] = reinterpret_cast<const uint4*>(lo)[v];
          }
          __threadfence();
          fence_proxy_async_all();
          epi_bar();
          if (etid == 0) {
            __threadfence();
            atomicAdd(p.flag + I.g, 1u);
          }
        }
      } else {
        // ---- C item: K^T W16 over the group's rows, one centre per thread; a chain = one row block ----
        const int n_chains = (I.st1 - I.st0) >> 1;
        if (lane == 0) wait_flag(p.flag + I.g, static_cast<unsigned int>(n_chains));   // wscale of the group is visible
        __syncwarp();
        // the group's scales: one L2 round trip per item (one word per thread), then broadcast reads from shared memory
        epi_bar();                                        // the previous item's readers of the staging words are done
        if (etid < 32 * n_chains) hand[4 + etid] = __float_as_uint(__ldcg(p.wscale + static_cast<int64_t>(I.rb) * 32 + etid));
        epi_bar();
        float acc[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) acc[t] = 0.f;
        for (int c = 0; c < n_chains; ++c) {
          float val[32];
          drain(val);
          const float4* sc = reinterpret_cast<const float4*>(hand + 4 + 32 * c);
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 s4 = sc[v];
            acc[4 * v + 0] = fmaf(val[4 * v + 0], s4.x, acc[4 * v + 0]); acc[4 * v + 1] = fmaf(val[4 * v + 1], s4.y, acc[4 * v + 1]);
            acc[4 * v + 2] = fmaf(val[4 * v + 2], s4.z, acc[4 * v + 2]); acc[4 * v + 3] = fmaf(val[4 * v + 3], s4.w, acc[4 * v + 3]);
          }
        }
        const int m = I.x * 128 + row;
        if (m < p.M) {
          float* orow = p.out + (static_cast<int64_t>(I.g) * p.M + m) * p.T_pad;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            if (4 * v < p.T_pad)
              *reinterpret_cast<float4*>(orow + 4 * v) = make_float4(acc[4 * v + 0], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, QTM_COLS);
}

int sweep_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

// groups of RG row blocks = partial slabs of the fused sweep
int panel16_sweep_slabs(int64_t n_rows) { return static_cast<int>((((n_rows + 127) / 128) + RG - 1) / RG); }

// workspace: [cnt n_rb][flag n_groups] (zeroed by every launch) [wscale n_rb x 32][A partial ring]
static size_t sweep_counter_bytes(int64_t n_rb, int64_t n_groups) { return static_cast<size_t>(round_up((n_rb + n_groups) * 4, 256)); }
size_t panel16_sweep_work_bytes(int64_t n_rows, int64_t M) {
  const int64_t n_rb = (n_rows + 127) / 128, n_groups = (n_rb + RG - 1) / RG;
  const int64_t n_ct = (M + 127) / 128, n_cs = (2 * n_ct + SW_STAGES - 1) / SW_STAGES;
  return sweep_counter_bytes(n_rb, n_groups) + static_cast<size_t>(round_up(n_rb * 32 * 4, 256)) +
         static_cast<size_t>(SW_RING) * n_cs * (RG * 128) * 32 * 4;
}

// out_partial[g] = K[rows of group g]^T (K[rows of group g] V) for every group of the chunk, from its resident panel P16;
// V16 / absmax_v as for launch_panel16_mmv; W16: [round_up(n_rows, 128) x 64] fp16 scratch (written here).
int launch_panel16_sweep(const void* P16, int64_t n_rows, int64_t M, const void* V16, const uint32_t* absmax_v, int T_pad,
                         void* W16, void* work, size_t work_bytes, float* out_partial, int n_slabs, cudaStream_t st) {
  if (n_rows <= 0 || M <= 0 || (T_pad != 16 && T_pad != 32) || (reinterpret_cast<uintptr_t>(P16) & 127) != 0 ||
      (reinterpret_cast<uintptr_t>(V16) & 127) != 0 || (reinterpret_cast<uintptr_t>(W16) & 127) != 0 ||
      (reinterpret_cast<uintptr_t>(work) & 255) != 0 || (reinterpret_cast<uintptr_t>(out_partial) & 15) != 0 || !absmax_v ||
      (reinterpret_cast<uintptr_t>(absmax_v) & 15) != 0)
    return set_error(ODF_ERR_ARG, "panel16_sweep: bad shape or alignment (P16, V16, W16 128-byte, work 256-byte aligned; T_pad 16 or 32)");
  if (n_slabs != panel16_sweep_slabs(n_rows)) return set_error(ODF_ERR_ARG, "panel16_sweep: n_slabs must come from odf_panel16_sweep_slabs");
  if (work_bytes < panel16_sweep_work_bytes(n_rows, M)) return set_error(ODF_ERR_ARG, "panel16_sweep: workspace too small");
  static DeviceOnce attr_once;
  bool& attr_set = attr_once.here();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(panel16_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(panel16_sweep_kernel)");
    attr_set = true;
  }
  SweepParams p;
  p.n_ct = static_cast<int>((M + 127) / 128);
  p.n_rb = static_cast<int>((n_rows + 127) / 128);
  p.n_groups = (p.n_rb + RG - 1) / RG;
  p.n_cstages = 2 * p.n_ct;
  p.n_cs = (p.n_cstages + SW_STAGES - 1) / SW_STAGES;
  p.lag = sweep_env("ODF_SWEEP_LAG", 2);
  if (p.lag < 0) p.lag = 0;
  if (p.lag > SW_RING - 2) p.lag = SW_RING - 2;
  p.ipb = RG * p.n_cs + p.n_ct;
  p.n_rows = static_cast<int>(n_rows);
  p.M = static_cast<int>(M);
  p.T_pad = T_pad;
  const int64_t plane_rows = static_cast<int64_t>(p.n_ct) * p.n_rb * 16;
  if (2 * plane_rows > 0x7fffffffll) return set_error(ODF_ERR_ARG, "panel16_sweep: chunk too large for 32-bit TMA coordinates");
  if (static_cast<int64_t>(p.n_groups + p.lag) * p.ipb > 0x7fffffffll) return set_error(ODF_ERR_ARG, "panel16_sweep: too many items");
  p.plane_rows = static_cast<int>(plane_rows);
  p.plane_elems = static_cast<int64_t>(p.n_ct) * p.n_rb * 16384;
  p.policy_a = sweep_env("ODF_SWEEP_POLICY_A", 1);
  p.policy_c = sweep_env("ODF_SWEEP_POLICY_C", 2);
  p.dbg_ncs = sweep_env("ODF_SWEEP_DEBUG_NCS", 0);
  p.P = static_cast<const __half*>(P16);
  p.absmax_v = absmax_v;
  uint8_t* wk = static_cast<uint8_t*>(work);
  const size_t cbytes = sweep_counter_bytes(p.n_rb, p.n_groups);
  p.cnt = reinterpret_cast<unsigned int*>(wk);
  p.flag = p.cnt + p.n_rb;
  p.wscale = reinterpret_cast<float*>(wk + cbytes);
  p.apart = reinterpret_cast<float*>(wk + cbytes + round_up(static_cast<int64_t>(p.n_rb) * 32 * 4, 256));
  p.W16 = static_cast<__half*>(W16);
  p.out = out_partial;
  cudaError_t e = cudaMemsetAsync(wk, 0, cbytes, st);
  if (e != cudaSuccess) return set_cuda_error(e, "panel16_sweep: memset");
  CUtensorMap tmP, tmV, tmW;
  int rc;
  if ((rc = make_map_plain_f16(&tmP, P16, 2 * plane_rows, 1024, 1024, 16, 256))) return rc;
  if ((rc = make_map_sw128(&tmV, V16, static_cast<int64_t>(p.n_ct) * 128, 64, 64, QR, 2))) return rc;
  if ((rc = make_map_sw128(&tmW, W16, static_cast<int64_t>(p.n_rb) * 128, 64, 64, QR, 2))) return rc;
  const int64_t n_items = static_cast<int64_t>(p.n_groups + p.lag) * p.ipb;
  const int sms = device_sm_count();
  const int grid = n_items < sms ? static_cast<int>(n_items) : sms;
  panel16_sweep_kernel<<<grid, 256, SW_SMEM, st>>>(tmP, tmV, tmW, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "panel16_sweep_kernel launch");
  return ODF_OK;
}

}  // namespace odf
