// CTA-pair version of the fused Gaussian-kernel tile (see odf_gauss_tile.cu for the operator, the operand format
// and the numerics): two CTAs of a cluster run every tcgen05.mma as one cta_group::2 instruction of M = 256.
// Each CTA keeps its own 128 rows (A tiles, TMEM accumulators, epilogue) but loads only HALF of every column tile
// (64 of the 128 centres of Q_hi / Q_lo, half of the V^T tile); the tensor cores read the other half from the peer's
// shared memory.  The single-CTA tile is bound by the L2 -> shared-memory operand feed (64 KB per k-block for 12 MMAs;
// measured ~16 TB/s chip-wide with the tensor pipe 2/3 busy): the pair moves 48 KB per CTA for the same MMA work, and
// the freed shared memory gives a fourth pipeline stage.
//
// Second contraction (K.V) in this kernel: the epilogue packs K as fp16 pairs hi = rn16(K), lo = rn16((K - hi) 2^12)
// IN PLACE over the fp32 S accumulator (32 S columns -> 16 columns of hi pairs + 16 of lo pairs, one tcgen05.st) and
// the tensor core contracts them straight from TMEM (kind::f16, K = 16 per MMA) with the fp16 hi/lo split of V^T
// (odf_split_rhs16: per-column power-of-two scales) into three accumulators W_hh, W_lh, W_hl that the read-out
// combines in fp32.  Compared with the tf32 hi/lo pair of the single-CTA kernel: no separate K_lo TMEM region (so the
// epilogue of tile n+1 never waits for the contraction of tile n), half the TMEM store traffic, 24 instead of 48
// MMAs per tile, and the same 2 x 11-bit significands.  The squared norms of the tile's centres arrive through a
// small shared-memory ring (bulk copies issued three tiles ahead by the otherwise idle allocator warp): as __ldg from
// the epilogue they missed the 26 KB of L1 left beside 200 KB of shared memory and waited on an L2 saturated by TMA.
//
// Only the K.V contraction mode (MODE_MMV, optional fp16-plane spill) is provided here; K_MM (store epilogue) and small
// problems stay on the single-CTA kernel.  Reference call sites are those of odf_gauss_tile.cu (falkon
// GaussianKernel.mmv / dmmv behind FALKONWrapper_with_centers_selection_incore.py:68,75-82).
//
// (4 + EPI_WARPS) warps per CTA: warp 0 operand producer, 1 MMA issuer (leader), 2 TMEM allocator + centre-norm ring
// producer, 3 V^T producer, 4.. epilogue (EPI_WARPS = 8 puts two warps on each TMEM lane quarter, each taking half of a
// tile's column chunks; measured no faster than 4 once the centre norms came from shared memory).
//
// Protocol (barriers at the same shared-memory offset in both CTAs):
//   FULL[s]   leader's copy only, count 2: each CTA's producer arrives with its own byte count, its TMA loads
//             complete_tx on the leader's barrier (cp.async.bulk.tensor ... cta_group::2)
//   EMPTY[s], SFULL[b], VEMPTY, WFULL: one arrival per CTA from the leader's multicast tcgen05.commit
//   PREADY, WEMPTY: leader's copy only, count 2 x 32 x EPI_WARPS: the epilogue threads of both CTAs arrive (remote for rank 1)
//   VFULL     leader's copy only, count 2 (V^T halves)
//   QFULL[4], QEMPTY[4]: per CTA, the ring of centre norms (producer: warp 2, consumers: the 4 epilogue warps)
#include <cstdlib>
#include <cuda_fp16.h>
#include "odf_ptx.cuh"
#include "odf_internal.h"

namespace odf {

namespace {

constexpr int BM = 128;                    // rows per CTA (TMEM lanes)
constexpr int BN = 128;                    // columns per tile
constexpr int BNH = BN / 2;                // columns of a tile held by one CTA
constexpr int BKB = 128;                   // bytes per k-block row
constexpr int NS = 4;                      // operand pipeline depth
constexpr int RT_BYTES = BM * BKB;         // 16 KB: one row-operand tile
constexpr int QT_BYTES = BNH * BKB;        // 8 KB: half a column-operand tile
constexpr int STAGE_BYTES = 2 * RT_BYTES + 2 * QT_BYTES;   // 48 KB
constexpr int MAX_TPAD = 32;
constexpr int V_ATOM_BYTES_MAX = (MAX_TPAD / 2) * 128;     // one [T_pad/2 x 64 fp16] box
constexpr int V_BYTES = 2 * 2 * V_ATOM_BYTES_MAX;          // hi+lo, 2 atoms each: 8 KB
constexpr int EPI_WARPS = 4;               // epilogue warps: 4 (one per TMEM lane quarter) or 8 (two; measured no faster)
constexpr int NTHREADS = (4 + EPI_WARPS) * 32;
constexpr int QN_SLOTS = 4;
constexpr int QN_BYTES = QN_SLOTS * BN * 4;                // ring of |c|^2 for 4 column tiles
constexpr int NUM_BARS = 2 * NS + 8 + 2 * QN_SLOTS;
constexpr int SMEM_BYTES = NS * STAGE_BYTES + V_BYTES + QN_BYTES + NUM_BARS * 8 + 16 + 1024;

// TMEM: S / packed K buffers 0 and 1, then the three K.V accumulators (hi.hi, lo.hi, hi.lo), 32 columns each
constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_W = 256, TM_COLS = 512;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct WorkItem {
  int row0;     // first row of the 256-row pair block
  int jt0, jt1; // column tiles [jt0, jt1)
  int split;
};

__device__ __forceinline__ WorkItem decode_item(const TileParams& p, int idx) {
  // same order as the single-CTA kernel, in units of pair blocks: groups of `group_rows` pair blocks x all splits
  const int per_full = p.group_rows * p.n_splits;
  const int n_full = p.n_rowblocks / p.group_rows;
  int g = idx / per_full;
  int rem, gsize;
  if (g < n_full) {
    rem = idx - g * per_full;
    gsize = p.group_rows;
  } else {
    g = n_full;
    rem = idx - n_full * per_full;
    gsize = p.n_rowblocks - n_full * p.group_rows;
  }
  const int split = rem / gsize;
  const int r = rem - split * gsize;
  WorkItem w;
  w.row0 = (g * p.group_rows + r) * (2 * BM);
  w.split = split;
  w.jt0 = split * p.tiles_per_split;
  w.jt1 = min(w.jt0 + p.tiles_per_split, p.n_coltiles);
  return w;
}

// instruction descriptors for M = 256 (cta_group::2)
__device__ __forceinline__ uint32_t idesc2_f16(uint32_t N) { return (1u << 4) | ((N >> 3) << 17) | ((256u >> 4) << 24); }
__device__ __forceinline__ uint32_t idesc2_tf32(uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((256u >> 4) << 24);
}
// D[tmem] (+)= A[tmem, packed fp16 pairs] * B[smem]^T, kind::f16, cta_group::2
__device__ __forceinline__ void mma_f16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

}  // namespace

// In this kernel TileParams::n_rowblocks counts 256-row PAIR blocks; rb128 = number of 128-row blocks (panel layout).
template <int KIND, int SPILL16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gauss_tile2_kernel(const __grid_constant__ CUtensorMap tmRh, const __grid_constant__ CUtensorMap tmRl,
                   const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                   const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl,
                   const TileParams p, const int rb128) {
  extern __shared__ uint8_t smem_raw[];
  // the dynamic shared window starts at the same offset in both CTAs, so the aligned base does too
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* stage_base = smem;
  uint8_t* v_base = smem + NS * STAGE_BYTES;
  float* qn_base = reinterpret_cast<float*>(v_base + V_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_base + V_BYTES + QN_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int B_FULL = 0, B_EMPTY = NS, B_SFULL = 2 * NS, B_PREADY = 2 * NS + 2, B_VFULL = 2 * NS + 3,
            B_VEMPTY = 2 * NS + 4, B_WFULL = 2 * NS + 5, B_WEMPTY = 2 * NS + 6, B_QFULL = 2 * NS + 7,
            B_QEMPTY = 2 * NS + 7 + QN_SLOTS, B_PREADY1 = 2 * NS + 7 + 2 * QN_SLOTS;
  // one "epilogue done" barrier per S buffer: a single barrier lets an epilogue warp that runs a tile ahead arrive twice
  // in one phase (see odf_gauss_tile.cu); nothing in this kernel delays a warp by a whole tile today, but nothing forbids it
  auto PREADY = [&](uint32_t b) { return BAR(b ? B_PREADY1 : B_PREADY); };

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmRh);
    tma_prefetch_desc(&tmRl);
    tma_prefetch_desc(&tmQh);
    tma_prefetch_desc(&tmQl);
    tma_prefetch_desc(&tmVh);
    tma_prefetch_desc(&tmVl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(BAR(B_FULL + s), 2);
      mbar_init(BAR(B_EMPTY + s), 1);
    }
    mbar_init(BAR(B_SFULL + 0), 1);
    mbar_init(BAR(B_SFULL + 1), 1);
    mbar_init(BAR(B_PREADY), 2 * EPI_WARPS * 32);
    mbar_init(BAR(B_PREADY1), 2 * EPI_WARPS * 32);
    for (int s = 0; s < QN_SLOTS; ++s) {
      mbar_init(BAR(B_QFULL + s), 1);
      mbar_init(BAR(B_QEMPTY + s), EPI_WARPS);
    }
    mbar_init(BAR(B_VFULL), 2);
    mbar_init(BAR(B_VEMPTY), 1);
    mbar_init(BAR(B_WFULL), 1);
    mbar_init(BAR(B_WEMPTY), 2 * EPI_WARPS * 32);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(smem_u32(tmem_slot), TM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // both CTAs' barriers are initialised before any remote arrival
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = p.n_rowblocks * p.n_splits;
  const int KB = p.kblocks;
  const int T_pad = p.T_pad;
  const int T_half = T_pad >> 1;
  constexpr int BK = (KIND == KIND_F16) ? 64 : 32;

  if (warp == 0) {
    // ======================= operand TMA producer (both CTAs) =======================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = pair; it < n_items; it += n_pairs) {
        const WorkItem w = decode_item(p, it);
        const int my_row0 = w.row0 + static_cast<int>(rank) * BM;
        for (int j = w.jt0; j < w.jt1; ++j) {
          const int my_col0 = j * BN + static_cast<int>(rank) * BNH;
          for (int kb = -1; kb < KB; ++kb) {
            mbar_wait(BAR(B_EMPTY + stage), phase ^ 1);
            const uint32_t full = BAR(B_FULL + stage);
            const uint32_t dst = smem_u32(stage_base) + stage * STAGE_BYTES;
            if (kb < 0) {
              // seed block: rows bring [-|x|^2/2 hi, lo, 0...], columns bring [1, 1, 0...]
              mbar_arrive_expect_tx_leader(full, RT_BYTES + QT_BYTES);
              tma_load_2d_pair(dst, &tmRh, full, KB * BK, my_row0);
              tma_load_2d_pair_hint(dst + 2 * RT_BYTES, &tmQh, full, (KB + 1) * BK, my_col0, kEvictLast);
            } else {
              mbar_arrive_expect_tx_leader(full, STAGE_BYTES);
              tma_load_2d_pair(dst, &tmRh, full, kb * BK, my_row0);
              tma_load_2d_pair(dst + RT_BYTES, &tmRl, full, kb * BK, my_row0);
              // the column operands (centres) are re-read by every row block of the launch: keep them in L2 against
              // the panel planes streaming through it
              tma_load_2d_pair_hint(dst + 2 * RT_BYTES, &tmQh, full, kb * BK, my_col0, kEvictLast);
              tma_load_2d_pair_hint(dst + 2 * RT_BYTES + QT_BYTES, &tmQl, full, kb * BK, my_col0, kEvictLast);
            }
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ======================= V^T half-tile TMA producer (both CTAs) =======================
    if (elect_one()) {
      uint32_t n = 0;
      const uint32_t atom_bytes = T_half * 128;
      const uint32_t full = BAR(B_VFULL);
      const uint32_t dst = smem_u32(v_base);
      const int vrow = static_cast<int>(rank) * T_half;
      for (int it = pair; it < n_items; it += n_pairs) {
        const WorkItem w = decode_item(p, it);
        for (int j = w.jt0; j < w.jt1; ++j, ++n) {
          mbar_wait(BAR(B_VEMPTY), (n & 1) ^ 1);
          mbar_arrive_expect_tx_leader(full, 4 * atom_bytes);
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            tma_load_2d_pair(dst + a * atom_bytes, &tmVh, full, j * BN + a * 64, vrow);
            tma_load_2d_pair(dst + (2 + a) * atom_bytes, &tmVl, full, j * BN + a * 64, vrow);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && leader) {
    // ======================= MMA issuer (leader CTA only, for both) =======================
    if (elect_one()) {
      const uint32_t idesc_s = (KIND == KIND_F16) ? idesc2_f16(BN) : idesc2_tf32(BN);
      const uint32_t idesc_pv = idesc2_f16(T_pad);
      const uint32_t atom_bytes = T_half * 128;
      const uint32_t sdesc_stage0 = (smem_u32(stage_base) & 0x3FFFFu) >> 4;
      const uint32_t sdesc_v = (smem_u32(v_base) & 0x3FFFFu) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t n = 0;
      uint32_t item_cnt = 0;
      bool pend = false, pend_first = false, pend_last = false;

      auto finish_prev = [&](uint32_t tile) {
        mbar_wait_cluster(PREADY(tile & 1), (tile >> 1) & 1);
        mbar_wait_cluster(BAR(B_VFULL), tile & 1);
        if (pend_first) {
          mbar_wait_cluster(BAR(B_WEMPTY), (item_cnt & 1) ^ 1);
          ++item_cnt;
        }
        tc_fence_after();
        const uint32_t t_k = tmem_base + ((tile & 1) ? TM_S1 : TM_S0);
        const uint32_t t_w = tmem_base + TM_W;
        if (!(p.dbg & 8))
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) {
          // 16 columns of K per k-step: chunk ks>>1 of the S buffer holds [16 columns of hi pairs | 16 of lo pairs]
          const uint32_t a_hi = t_k + 32 * (ks >> 1) + 8 * (ks & 1);
          const uint32_t a_lo = a_hi + 16;
          const uint32_t off = (ks >> 2) * atom_bytes + (ks & 3) * 32;
          const uint64_t b_hi = kSdescSw128Hi | static_cast<uint64_t>(sdesc_v + (off >> 4));
          const uint64_t b_lo = kSdescSw128Hi | static_cast<uint64_t>(sdesc_v + ((2 * atom_bytes + off) >> 4));
          const uint32_t acc = (pend_first && ks == 0) ? 0u : 1u;
          mma_f16_ts_pair(t_w, a_hi, b_hi, idesc_pv, acc);
          mma_f16_ts_pair(t_w + 32, a_lo, b_hi, idesc_pv, acc);
          mma_f16_ts_pair(t_w + 64, a_hi, b_lo, idesc_pv, acc);
        }
        tc_commit_pair(BAR(B_VEMPTY));
        if (pend_last) tc_commit_pair(BAR(B_WFULL));
      };

      bool ready = false;
      for (int it = pair; it < n_items; it += n_pairs) {
        const WorkItem w = decode_item(p, it);
        const bool last_item = (it + n_pairs >= n_items);
        for (int j = w.jt0; j < w.jt1; ++j) {
          const uint32_t t_s = tmem_base + ((n & 1) ? TM_S1 : TM_S0);
          for (int kb = -1; kb < KB; ++kb) {
            const uint32_t sd = sdesc_stage0 + stage * (STAGE_BYTES >> 4);
            int nstage = stage + 1;
            uint32_t nphase = phase;
            if (nstage == NS) { nstage = 0; nphase ^= 1; }
            const bool more = !(last_item && j == w.jt1 - 1 && kb == KB - 1);
            if (!ready) mbar_wait_cluster(BAR(B_FULL + stage), phase);
            tc_fence_after();
            auto step = [&](int ks) {               // one 32-byte k-step: lo.hi + hi.lo + hi.hi
              const uint64_t a_hi = kSdescSw128Hi | static_cast<uint64_t>(sd + ((ks * 32) >> 4));
              const uint64_t a_lo = kSdescSw128Hi | static_cast<uint64_t>(sd + ((RT_BYTES + ks * 32) >> 4));
              const uint64_t b_hi = kSdescSw128Hi | static_cast<uint64_t>(sd + ((2 * RT_BYTES + ks * 32) >> 4));
              const uint64_t b_lo = kSdescSw128Hi | static_cast<uint64_t>(sd + ((2 * RT_BYTES + QT_BYTES + ks * 32) >> 4));
              mma_ss_pair<KIND>(t_s, a_lo, b_hi, idesc_s, 1u);
              mma_ss_pair<KIND>(t_s, a_hi, b_lo, idesc_s, 1u);
              mma_ss_pair<KIND>(t_s, a_hi, b_hi, idesc_s, 1u);
            };
            if (p.dbg & 2) {
            } else if (kb < 0) {
              mma_ss_pair<KIND>(t_s, kSdescSw128Hi | static_cast<uint64_t>(sd),
                                kSdescSw128Hi | static_cast<uint64_t>(sd + ((2 * RT_BYTES) >> 4)), idesc_s, 0u);
            } else {
              step(0); step(1); step(2);
            }
            ready = more && mbar_test_wait_cluster(BAR(B_FULL + nstage), nphase) != 0;
            if (kb >= 0 && !(p.dbg & 2)) step(3);
            tc_commit_pair(BAR(B_EMPTY + stage));
            stage = nstage;
            phase = nphase;
          }
          tc_commit_pair(BAR(B_SFULL + (n & 1)));
          if (pend) finish_prev(n - 1);
          pend = true;
          pend_first = (j == w.jt0);
          pend_last = (j == w.jt1 - 1);
          ++n;
        }
      }
      if (pend) finish_prev(n - 1);
    }
    __syncwarp();
  } else if (warp == 2) {
    // ======================= centre-norm ring producer (both CTAs) =======================
    if (elect_one()) {
      uint32_t n = 0;
      for (int it = pair; it < n_items; it += n_pairs) {
        const WorkItem w = decode_item(p, it);
        for (int j = w.jt0; j < w.jt1; ++j, ++n) {
          const uint32_t slot = n & (QN_SLOTS - 1);
          mbar_wait(BAR(B_QEMPTY + slot), ((n / QN_SLOTS) & 1) ^ 1);
          mbar_arrive_expect_tx(BAR(B_QFULL + slot), BN * 4);
          bulk_g2s(smem_u32(qn_base + slot * BN), p.qnorm + static_cast<int64_t>(j) * BN, BN * 4, BAR(B_QFULL + slot));
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ======================= epilogue (both CTAs, own 128 rows) =======================
    // with EPI_WARPS = 8: warps 4-7 take the column chunks 0-1 of every tile, warps 8-11 the chunks 2-3
    const int q = warp & 3;
    const int set = (warp - 4) >> 2;
    constexpr int CH_PER_SET = (BN / 32) / (EPI_WARPS / 4);
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float nsl2 = p.neg_scale_log2;
    const float s_r = __ldg(p.r_scale), s_q = __ldg(p.q_scale);
    const float m2inv = -2.f / (s_r * s_q);
    const float one_m_rho = 1.f - s_r / s_q;
    uint32_t n = 0;
    uint32_t item_cnt = 0;
    for (int it = pair; it < n_items; it += n_pairs, ++item_cnt) {
      const WorkItem w = decode_item(p, it);
      const int my_row0 = w.row0 + static_cast<int>(rank) * BM;
      const int grow = my_row0 + row;
      const float rn = ((grow < p.n_rows) ? __ldg(p.rnorm + grow) : 0.f) * one_m_rho;
      const int my_rb = my_row0 >> 7;
      for (int j = w.jt0; j < w.jt1; ++j, ++n) {
        const uint32_t b = n & 1;
        mbar_wait_warp(BAR(B_SFULL + b), (n >> 1) & 1);
        tc_fence_after();
        const uint32_t t_s = tmem_base + lane_off + (b ? TM_S1 : TM_S0);
        const uint32_t slot = n & (QN_SLOTS - 1);
        mbar_wait_warp(BAR(B_QFULL + slot), (n / QN_SLOTS) & 1);
        const float4* qs = reinterpret_cast<const float4*>(qn_base + slot * BN);
#pragma unroll 1
        for (int ch = set * CH_PER_SET; ch < (set + 1) * CH_PER_SET; ++ch) {
          uint32_t s[32];
          __syncwarp();
          tmem_ld32(t_s + ch * 32, s);
          float qn[32];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 t = (p.dbg & 32) ? make_float4(400.f, 400.f, 400.f, 400.f) : qs[ch * 8 + v];   // warp-wide broadcast
            qn[4 * v + 0] = t.x; qn[4 * v + 1] = t.y; qn[4 * v + 2] = t.z; qn[4 * v + 3] = t.w;
          }
          tc_wait_ld();
          if (p.dbg & 4) continue;
          uint32_t kp[32];                       // [0,16): hi pairs, [16,32): lo pairs (columns 2i, 2i+1 of the chunk)
          uint32_t pl[SPILL16 ? 8 : 1];          // the spilled lo plane: one byte per value (lo8_of)
          if (SPILL16) {
#pragma unroll
            for (int i = 0; i < 8; ++i) pl[i] = 0u;
          }
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float d0 = fmaf(m2inv, __uint_as_float(s[c]), rn + qn[c]);
            float d1 = fmaf(m2inv, __uint_as_float(s[c + 1]), rn + qn[c + 1]);
            d0 = fmaxf(d0, 0.f);
            d1 = fmaxf(d1, 0.f);
            const float k0 = ex2_approx(d0 * nsl2), k1 = ex2_approx(d1 * nsl2);
            const __half2 h = __floats2half2_rn(k0, k1);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn((k0 - hf.x) * 4096.f, (k1 - hf.y) * 4096.f);
            kp[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
            kp[16 + (c >> 1)] = *reinterpret_cast<const uint32_t*>(&l);
            if (SPILL16) pl[c >> 2] |= (lo8_of(k0 - hf.x) << (8 * (c & 3))) | (lo8_of(k1 - hf.y) << (8 * (c & 3) + 8));
          }
          tmem_st32(t_s + ch * 32, kp);          // in place: the chunk's 32 fp32 columns now hold its packed K
          if (SPILL16 && my_rb < rb128) {
            const int64_t e0 = (((static_cast<int64_t>(j) * rb128 + my_rb) * 16 + ch * 4) * 128 + row) * 8;
            __half* dst = p.panel16 + e0;
            uint8_t* dlo = reinterpret_cast<uint8_t*>(p.panel16 + p.panel16_plane) + e0;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              // st.global.cs: written once, read once by the panel kernel -> first in line for eviction
              __stcs(reinterpret_cast<uint4*>(dst + v * 1024), make_uint4(kp[4 * v], kp[4 * v + 1], kp[4 * v + 2], kp[4 * v + 3]));
              __stcs(reinterpret_cast<uint2*>(dlo + v * 1024), make_uint2(pl[2 * v], pl[2 * v + 1]));
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_QEMPTY + slot));
        tc_wait_st();
        tc_fence_before();
        mbar_arrive_leader(PREADY(b));
      }
      // W for this item is complete once the last tile's contraction has retired.
      mbar_wait_warp(BAR(B_WFULL), item_cnt & 1);
      tc_fence_after();
      const uint32_t t_w = tmem_base + lane_off + TM_W;
      float* orow = p.out + static_cast<int64_t>(w.split) * p.split_stride + static_cast<int64_t>(grow) * T_pad;
      for (int c0 = 16 * set; c0 < T_pad; c0 += 16 * (EPI_WARPS / 4)) {
        uint32_t r[16];
        float acc[16];
        __syncwarp();
        tmem_ld16(t_w + 32 + c0, r);              // lo.hi  (x 2^-12)
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(r[i]) * (1.f / 4096.f);
        tmem_ld16(t_w + 64 + c0, r);              // hi.lo  (x 2^-11)
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(__uint_as_float(r[i]), 1.f / 2048.f, acc[i]);
        tmem_ld16(t_w + c0, r);                   // hi.hi
        tc_wait_ld();
        if (grow < p.n_rows) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 mx = __ldg(reinterpret_cast<const uint4*>(p.v_absmax + c0) + v);   // per-column scales of V16
            float4 t;
            t.x = (__uint_as_float(r[4 * v + 0]) + acc[4 * v + 0]) * w16_scale_from_bits(mx.x, true);
            t.y = (__uint_as_float(r[4 * v + 1]) + acc[4 * v + 1]) * w16_scale_from_bits(mx.y, true);
            t.z = (__uint_as_float(r[4 * v + 2]) + acc[4 * v + 2]) * w16_scale_from_bits(mx.z, true);
            t.w = (__uint_as_float(r[4 * v + 3]) + acc[4 * v + 3]) * w16_scale_from_bits(mx.w, true);
            *reinterpret_cast<float4*>(orow + c0 + 4 * v) = t;
          }
        }
      }
      tc_fence_before();
      mbar_arrive_leader(BAR(B_WEMPTY));
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // the peer may still be signalling this CTA's barriers / reading its smem
  if (warp == 2) tmem_dealloc_pair(tmem_base, TM_COLS);
}

// ----------------------------------------------------------------------------- host side
namespace {
int g2_num_sms() { return device_sm_count(); }
}  // namespace

// The pair kernel serves MODE_MMV launches with enough rows to keep every pair busy.
bool tile2_rows_eligible(int64_t n_rows) {
  const char* e = getenv("ODF_TILE_PAIR");
  if (e && atoi(e) == 0) return false;
  return n_rows >= 2 * BM * 32;
}
bool tile2_eligible(const TileLaunch& L) {
  return L.mode == MODE_MMV && L.panel == nullptr && L.vt16_hi != nullptr && tile2_rows_eligible(L.n_rows);
}

int launch_gauss_tile2(const TileLaunch& L, cudaStream_t stream) {
  if (L.kind != KIND_TF32 && L.kind != KIND_F16) return set_error(ODF_ERR_ARG, "unknown operand kind");
  const int64_t BK = kblock_elems(L.kind);
  const int esize = L.kind == KIND_F16 ? 2 : 4;
  const int64_t pitch = L.d_pad + 2 * BK;
  if (L.d_pad % BK != 0 || L.d_pad <= 0) return set_error(ODF_ERR_ARG, "d_pad must be a positive multiple of the k-block width");
  if (L.n_rows <= 0 || L.n_cols <= 0) return set_error(ODF_ERR_ARG, "empty operand");
  if (L.mode != MODE_MMV || !(L.T_pad == 16 || L.T_pad == 32)) return set_error(ODF_ERR_ARG, "pair tile: MODE_MMV with T_pad 16 or 32");
  static DeviceOnce attr_once;
  bool& attr_set = attr_once.here();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gauss_tile2_kernel<KIND_TF32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gauss_tile2_kernel<KIND_F16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gauss_tile2_kernel<KIND_TF32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gauss_tile2_kernel<KIND_F16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gauss_tile2_kernel)");
    attr_set = true;
  }
  CUtensorMap mRh, mRl, mQh, mQl, mVh, mVl;
  int rc;
  if ((rc = make_map_sw128(&mRh, L.r_hi, L.n_rows, pitch, pitch, BM, esize))) return rc;
  if ((rc = make_map_sw128(&mRl, L.r_lo, L.n_rows, pitch, pitch, BM, esize))) return rc;
  if ((rc = make_map_sw128(&mQh, L.q_hi, L.n_cols, pitch, pitch, BNH, esize))) return rc;
  if ((rc = make_map_sw128(&mQl, L.q_lo, L.n_cols, pitch, pitch, BNH, esize))) return rc;
  if (!L.vt16_hi || !L.vt16_lo || !L.v_absmax || L.ldvt16 % 64 != 0 || (reinterpret_cast<uintptr_t>(L.v_absmax) & 15) != 0)
    return set_error(ODF_ERR_ARG, "pair tile: needs the fp16 split of V^T (odf_split_rhs16) and its 16-byte aligned scales");
  if ((rc = make_map_sw128(&mVh, L.vt16_hi, L.T_pad, L.ldvt16, L.ldvt16, L.T_pad / 2, 2))) return rc;
  if ((rc = make_map_sw128(&mVl, L.vt16_lo, L.T_pad, L.ldvt16, L.ldvt16, L.T_pad / 2, 2))) return rc;
  TileParams p{};
  p.n_rows = static_cast<int>(L.n_rows);
  p.n_cols = static_cast<int>(L.n_cols);
  p.kblocks = static_cast<int>(L.d_pad / BK);
  p.T_pad = L.T_pad;
  p.mode = L.mode;
  const int rb128 = static_cast<int>((L.n_rows + BM - 1) / BM);
  p.n_rowblocks = static_cast<int>((L.n_rows + 2 * BM - 1) / (2 * BM));      // pair blocks
  p.n_coltiles = static_cast<int>((L.n_cols + BN - 1) / BN);
  int splits = L.n_splits > 0 ? L.n_splits : 1;
  if (splits > p.n_coltiles) splits = p.n_coltiles;
  p.tiles_per_split = (p.n_coltiles + splits - 1) / splits;
  p.n_splits = (p.n_coltiles + p.tiles_per_split - 1) / p.tiles_per_split;
  if (p.n_splits != L.n_splits) return set_error(ODF_ERR_ARG, "n_splits must divide the column tiles without empty splits (use odf_tile_splits)");
  const int pairs_max = g2_num_sms() / 2;
  p.group_rows = pairs_max / p.n_splits;
  if (p.group_rows < 1) p.group_rows = 1;
  p.neg_scale_log2 = static_cast<float>(-1.4426950408889634 / (2.0 * double(L.sigma) * double(L.sigma)));
  p.rnorm = L.r_norm;
  p.qnorm = L.q_norm;
  p.r_scale = L.r_scale;
  p.q_scale = L.q_scale;
  p.out = L.out;
  p.ldo = L.ldo;
  p.split_stride = L.split_stride;
  p.panel = nullptr;
  p.ldpanel = 0;
  p.v_absmax = L.v_absmax;
  p.panel16 = static_cast<__half*>(L.panel16);
  p.panel16_plane = static_cast<int64_t>(p.n_coltiles) * rb128 * 16384;
  if (L.panel16 != nullptr && (reinterpret_cast<uintptr_t>(L.panel16) & 127) != 0)
    return set_error(ODF_ERR_ARG, "panel16 must be 128-byte aligned");
  {
    const char* e = getenv("ODF_TILE_DEBUG");     // timing experiments: 2 = no S MMAs, 4 = no epilogue math, 8 = no K.V MMAs
    p.dbg = e ? atoi(e) : 0;
  }
  p.store_vec4 = 0;
  const int n_items = p.n_rowblocks * p.n_splits;
  const int n_pairs = n_items < pairs_max ? n_items : pairs_max;
  const int grid = 2 * n_pairs;
  if (L.kind == KIND_F16 && L.panel16 != nullptr)
    gauss_tile2_kernel<KIND_F16, 1><<<grid, NTHREADS, SMEM_BYTES, stream>>>(mRh, mRl, mQh, mQl, mVh, mVl, p, rb128);
  else if (L.kind == KIND_F16)
    gauss_tile2_kernel<KIND_F16, 0><<<grid, NTHREADS, SMEM_BYTES, stream>>>(mRh, mRl, mQh, mQl, mVh, mVl, p, rb128);
  else if (L.panel16 != nullptr)
    gauss_tile2_kernel<KIND_TF32, 1><<<grid, NTHREADS, SMEM_BYTES, stream>>>(mRh, mRl, mQh, mQl, mVh, mVl, p, rb128);
  else
    gauss_tile2_kernel<KIND_TF32, 0><<<grid, NTHREADS, SMEM_BYTES, stream>>>(mRh, mRl, mQh, mQl, mVh, mVl, p, rb128);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "gauss_tile2_kernel launch");
  return ODF_OK;
}

}  // namespace odf
