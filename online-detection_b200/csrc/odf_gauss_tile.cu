// Fused Gaussian-kernel tile for sm_100a (tcgen05 + TMEM + TMA).
//
// One persistent, warp-specialised kernel evaluates tiles of
//     K[r, q] = exp(-(|x_r|^2 + |c_q|^2 - 2 x_r.c_q) / (2 sigma^2))
// for a block of 128 "row" points against a stream of 128-wide "column" point tiles and, without
// ever writing K to memory, contracts each tile with a block of right-hand sides:
//     W[r, :] += K[r, tile] . V[tile, :]
// This is the operator behind every hot call of the reference path:
//   * predict / minibootstrap scoring  falkon `GaussianKernel.mmv`  (reference call sites:
//     src/modules/region-classifier/FALKONWrapper_with_centers_selection_incore.py:75-82,
//     .../box_head/roi_box_predictors.py:136,158, .../mask_head/roi_mask_predictors.py:61,90,
//     .../rpn/rpn.py:197,225)
//   * the CG operator K_nm^T (K_nm V) of `GaussianKernel.dmmv` (via FALKONWrapper.train ->
//     InCoreFalkon.fit, same file :58-68): first pass rows = data, columns = centres; second pass
//     rows = centres, columns = data (the Gaussian kernel is symmetric, K(X,C)^T = K(C,X)).
//   * K_MM for the preconditioner (store epilogue, MODE_STORE).
//
// x.c runs on the tensor cores as a 3-pass split product: operands are pre-split (odf_vec.cu) into
// hi = rn11(x), lo = rn11(x - hi) (11-bit significands) and the tile accumulates
// hi.hi + lo.hi + hi.lo in fp32 TMEM, which restores fp32-grade distances.  Two operand kinds share
// the kernel (template): KIND_TF32 (hi/lo stored as tf32 in fp32 words, kind::tf32 MMAs, "3xTF32")
// and KIND_F16 (hi/lo stored as fp16 after a power-of-two scaling of the point set, kind::f16
// MMAs at twice the tf32 rate and half the operand bytes; same 2 x 11 significand bits).
// The tensor core accumulates in fp32 with truncation, so a long chain of same-sign partial sums
// drifts; every tile therefore starts from a rank-1 MMA that seeds the accumulator with
// -|x_r|^2/2 (an extra k-block appended to the operand arrays by the pre-pass), which keeps the
// running sums of near-neighbour pairs centred on zero: acc = x.c - |x|^2/2, D = |c|^2 - 2 acc.
// The epilogue clamps D at 0, applies exp2 and splits K into tf32 hi/lo; K_hi overwrites the S
// accumulator in place and together with K_lo feeds the second tensor-core contraction straight
// from TMEM (A operand in TMEM, V^T tiles from shared memory), again as a 3xTF32 product.
//
// Warp roles (256 threads, 1 CTA / SM):
//   warp 0  lane 0 : TMA producer for the operand pipeline (R_hi, R_lo, Q_hi, Q_lo k-blocks)
//   warp 1  lane 0 : tcgen05.mma issuer (S tiles and the K.V contraction)
//   warp 2         : TMEM allocator
//   warp 3  lane 0 : TMA producer for the V^T tiles
//   warps 4..7     : epilogue (one TMEM lane == one row per thread)
#include <cstdlib>
#include <cuda_fp16.h>
#include "odf_ptx.cuh"
#include "odf_internal.h"

namespace odf {

namespace {

constexpr int BM = 128;                    // rows per tile (TMEM lanes)
constexpr int BN = 128;                    // columns per tile
constexpr int BKB = 128;                   // bytes per k-block row (= one 128B swizzle atom)
constexpr int NS = 3;                      // operand pipeline depth
constexpr int TILE_BYTES = BM * BKB;        // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;  // R_hi, R_lo, Q_hi, Q_lo
constexpr int MAX_TPAD = 32;
constexpr int V_ATOM_BYTES_MAX = MAX_TPAD * 128;         // one [T_pad x 32] box
constexpr int V_BYTES = 2 * 4 * V_ATOM_BYTES_MAX;        // hi+lo, 4 atoms each
constexpr int NUM_BARS = 2 * NS + 9;
constexpr int SMEM_BYTES = NS * STAGE_BYTES + V_BYTES + NUM_BARS * 8 + 16 + 1024;

// TMEM column map (512 columns allocated)
constexpr uint32_t TM_S0 = 0;      // S / K_hi buffer 0
constexpr uint32_t TM_S1 = 128;    // S / K_hi buffer 1
constexpr uint32_t TM_PLO = 256;   // K_lo
constexpr uint32_t TM_W = 384;     // W accumulator (T_pad columns)
constexpr uint32_t TM_COLS = 512;

struct WorkItem {
  int row0;     // first row of the block
  int jt0;      // first column tile
  int jt1;      // one past the last column tile
  int split;    // split index (selects the partial output slab)
};

__device__ __forceinline__ WorkItem decode_item(const TileParams& p, int idx) {
  // Items are ordered so that CTAs running side by side share operand tiles in L2:
  // groups of `group_rows` row blocks x all splits; inside a group split-major.
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
  w.row0 = (g * p.group_rows + r) * BM;
  w.split = split;
  w.jt0 = split * p.tiles_per_split;
  w.jt1 = min(w.jt0 + p.tiles_per_split, p.n_coltiles);
  return w;
}

}  // namespace

// SPILL16 = 1: the epilogue also writes every K tile as two planes (hi = rn16(K) fp16, lo = rni((K - hi) * 2^19) + 128 bytes)
// in the tile-blocked layout odf_panel16.cu streams back with TMA: [plane][column tile][row block][16 groups][128 rows][8].
// LINEAR = 1 (MODE_STORE only, EXPERIMENTAL): the split GEMM A B^T behind the blocked preconditioner build
// (odf/precond_blocked.py).  Operands come from odf_prepare_points_linear (zero seed block, so the accumulator is
// s_r s_q (x . c) exactly) and the epilogue stores lin_alpha (x . c) + lin_beta out instead of the Gaussian kernel.
template <int KIND, int SPILL16, int LINEAR = 0>
__global__ void __launch_bounds__(256, 1)
gauss_tile_kernel(const __grid_constant__ CUtensorMap tmRh, const __grid_constant__ CUtensorMap tmRl,
                  const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                  const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl,
                  const TileParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* stage_base = smem;
  uint8_t* v_base = smem + NS * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_base + V_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  // barrier indices
  const int B_FULL = 0, B_EMPTY = NS, B_SFULL = 2 * NS, B_PREADY = 2 * NS + 2,
            B_PVDONE = 2 * NS + 3, B_VFULL = 2 * NS + 4, B_VEMPTY = 2 * NS + 5,
            B_WFULL = 2 * NS + 6, B_WEMPTY = 2 * NS + 7, B_PREADY1 = 2 * NS + 8;
  // "epilogue done with S buffer b" has ONE BARRIER PER BUFFER: with a single barrier an epilogue warp that runs a tile
  // ahead of the others (rows past n_rows skip their stores; in MODE_STORE nothing else holds it back) arrives twice in
  // one phase, the phase completes before the slow warps have read their rows and the issuer overwrites the buffer --
  // wrong entries in the ragged last row block, or a dead-locked hand-shake (round-2 bring-up of the LINEAR variant).
  auto PREADY = [&](uint32_t b) { return BAR(b ? B_PREADY1 : B_PREADY); };

  // warp index through a shuffle: tells ptxas it is warp-uniform (role branches stay uniform)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const bool mode_mmv = (p.mode == MODE_MMV);
  long long dbg_c0 = 0;
  unsigned long long dbg_t0 = 0;
  if ((p.dbg & 16) && threadIdx.x == 0) {
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmRh);
    tma_prefetch_desc(&tmRl);
    tma_prefetch_desc(&tmQh);
    tma_prefetch_desc(&tmQl);
    if (mode_mmv) {
      tma_prefetch_desc(&tmVh);
      tma_prefetch_desc(&tmVl);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(BAR(B_FULL + s), 1);
      mbar_init(BAR(B_EMPTY + s), 1);
    }
    mbar_init(BAR(B_SFULL + 0), 1);
    mbar_init(BAR(B_SFULL + 1), 1);
    mbar_init(BAR(B_PREADY), 128);
    mbar_init(BAR(B_PREADY1), 128);
    mbar_init(BAR(B_PVDONE), 1);
    mbar_init(BAR(B_VFULL), 1);
    mbar_init(BAR(B_VEMPTY), 1);
    mbar_init(BAR(B_WFULL), 1);
    mbar_init(BAR(B_WEMPTY), 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = p.n_rowblocks * p.n_splits;
  const int KB = p.kblocks;
  const int T_pad = p.T_pad;
  constexpr int BK = (KIND == KIND_F16) ? 64 : 32;      // elements per k-block

  // Producer / issuer roles: ONE elected lane of the warp runs the whole role loop.  `elect_one()`
  // (elect.sync) rather than `lane == 0` lets ptxas keep descriptors and barrier addresses in uniform
  // registers; with `lane == 0` it wrapped every UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST retry
  // loop and the issuing thread, not the tensor pipe, set the pace (~100 cycles per 64-cycle MMA).
  if (warp == 0) {
    // ======================= operand TMA producer =======================
    if (elect_one()) {
      const bool no_tma = (p.dbg & 1) != 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const WorkItem w = decode_item(p, it);
        for (int j = w.jt0; j < w.jt1; ++j) {
          const int col0 = j * BN;
          for (int kb = -1; kb < KB; ++kb) {
            mbar_wait(BAR(B_EMPTY + stage), phase ^ 1);
            const uint32_t full = BAR(B_FULL + stage);
            const uint32_t dst = smem_u32(stage_base) + stage * STAGE_BYTES;
            if (no_tma) {
              mbar_arrive(full);
            } else if (kb < 0) {
              // seed block: rows bring [-|x|^2/2 hi, lo, 0...], columns bring [1, 1, 0...]
              mbar_arrive_expect_tx(full, 2 * TILE_BYTES);
              tma_load_2d(dst + 0 * TILE_BYTES, &tmRh, full, KB * BK, w.row0);
              tma_load_2d(dst + 2 * TILE_BYTES, &tmQh, full, (KB + 1) * BK, col0);
            } else {
              mbar_arrive_expect_tx(full, STAGE_BYTES);
              tma_load_2d(dst + 0 * TILE_BYTES, &tmRh, full, kb * BK, w.row0);
              tma_load_2d(dst + 1 * TILE_BYTES, &tmRl, full, kb * BK, w.row0);
              tma_load_2d(dst + 2 * TILE_BYTES, &tmQh, full, kb * BK, col0);
              tma_load_2d(dst + 3 * TILE_BYTES, &tmQl, full, kb * BK, col0);
            }
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 3 && mode_mmv && !(p.dbg & 64)) {
    // ======================= V^T tile TMA producer =======================
    if (elect_one()) {
      uint32_t n = 0;  // tile counter
      const uint32_t atom_bytes = T_pad * 128;
      const uint32_t full = BAR(B_VFULL);
      const uint32_t dst = smem_u32(v_base);
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const WorkItem w = decode_item(p, it);
        for (int j = w.jt0; j < w.jt1; ++j, ++n) {
          mbar_wait(BAR(B_VEMPTY), (n & 1) ^ 1);
          mbar_arrive_expect_tx(full, 8 * atom_bytes);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            tma_load_2d(dst + a * atom_bytes, &tmVh, full, j * BN + a * 32, 0);
            tma_load_2d(dst + (4 + a) * atom_bytes, &tmVl, full, j * BN + a * 32, 0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      const uint32_t idesc_s = (KIND == KIND_F16) ? make_idesc_f16(BM, BN) : make_idesc_tf32(BM, BN);
      const uint32_t idesc_pv = make_idesc_tf32(BM, T_pad);
      const uint32_t atom_bytes = T_pad * 128;
      const bool no_smma = (p.dbg & 2) != 0, no_pv = (p.dbg & 8) != 0;
      // descriptor low words: (address >> 4); tiles at fixed offsets add (offset >> 4)
      const uint32_t sdesc_stage0 = (smem_u32(stage_base) & 0x3FFFFu) >> 4;
      const uint32_t sdesc_v = (smem_u32(v_base) & 0x3FFFFu) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t n = 0;         // tiles whose S-MMAs have been issued
      uint32_t item_cnt = 0;  // items whose first PV has been issued
      bool pend = false, pend_first = false, pend_last = false;

      // second contraction (or, in MODE_STORE, just the hand-back of the S buffer) for tile n-1
      auto finish_prev = [&](uint32_t tile) {
        if (p.dbg & 64) return;                       // timing experiment: issuer + producer only
        mbar_wait(PREADY(tile & 1), (tile >> 1) & 1);
        if (mode_mmv) {
          mbar_wait(BAR(B_VFULL), tile & 1);
          if (pend_first) {
            mbar_wait(BAR(B_WEMPTY), (item_cnt & 1) ^ 1);
            ++item_cnt;
          }
          tc_fence_after();
          const uint32_t t_hi = tmem_base + ((tile & 1) ? TM_S1 : TM_S0);
          const uint32_t t_lo = tmem_base + TM_PLO;
          const uint32_t t_w = tmem_base + TM_W;
          if (!no_pv) {
#pragma unroll
            for (int ks = 0; ks < BN / 8; ++ks) {
              const uint32_t off = (ks >> 2) * atom_bytes + (ks & 3) * 32;
              const uint64_t b_hi = kSdescSw128Hi | static_cast<uint64_t>(sdesc_v + (off >> 4));
              const uint64_t b_lo = kSdescSw128Hi | static_cast<uint64_t>(sdesc_v + ((4 * atom_bytes + off) >> 4));
              mma_tf32_ts(t_w, t_hi + ks * 8, b_hi, idesc_pv, (pend_first && ks == 0) ? 0u : 1u);
              mma_tf32_ts(t_w, t_lo + ks * 8, b_hi, idesc_pv, 1u);
              mma_tf32_ts(t_w, t_hi + ks * 8, b_lo, idesc_pv, 1u);
            }
          }
          tc_commit(BAR(B_VEMPTY));
          tc_commit(BAR(B_PVDONE));
          if (pend_last) tc_commit(BAR(B_WFULL));
        }
      };

      // The probe for the NEXT k-block's stage sits between the MMAs of the current one: the tensor
      // pipe holds only a few queued MMAs, so a full-barrier poll + descriptor set-up at the k-block
      // boundary (~300 cycles of dependent scalar work) would otherwise drain it every 12 MMAs
      // (measured with tools/umma_probe.cu: 81 -> 72 cycles per 64-cycle MMA).
      bool ready = false;     // has the current k-block's stage already been seen full?
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const WorkItem w = decode_item(p, it);
        const bool last_item = (it + static_cast<int>(gridDim.x) >= n_items);
        for (int j = w.jt0; j < w.jt1; ++j) {
          const uint32_t t_s = tmem_base + ((n & 1) ? TM_S1 : TM_S0);
          for (int kb = -1; kb < KB; ++kb) {
            const uint32_t sd = sdesc_stage0 + stage * (STAGE_BYTES >> 4);
            int nstage = stage + 1;
            uint32_t nphase = phase;
            if (nstage == NS) { nstage = 0; nphase ^= 1; }
            const bool more = !(last_item && j == w.jt1 - 1 && kb == KB - 1);
            if (!ready) mbar_wait(BAR(B_FULL + stage), phase);
            tc_fence_after();
            auto step = [&](int ks) {               // one 32-byte k-step: lo.hi + hi.lo + hi.hi
              const uint64_t a_hi = kSdescSw128Hi | static_cast<uint64_t>(sd + ((0 * TILE_BYTES + ks * 32) >> 4));
              const uint64_t a_lo = kSdescSw128Hi | static_cast<uint64_t>(sd + ((1 * TILE_BYTES + ks * 32) >> 4));
              const uint64_t b_hi = kSdescSw128Hi | static_cast<uint64_t>(sd + ((2 * TILE_BYTES + ks * 32) >> 4));
              const uint64_t b_lo = kSdescSw128Hi | static_cast<uint64_t>(sd + ((3 * TILE_BYTES + ks * 32) >> 4));
              mma_ss<KIND>(t_s, a_lo, b_hi, idesc_s, 1u);      // small terms first
              mma_ss<KIND>(t_s, a_hi, b_lo, idesc_s, 1u);
              mma_ss<KIND>(t_s, a_hi, b_hi, idesc_s, 1u);
            };
            if (no_smma) {
            } else if (kb < 0) {
              // rank-1 seed: acc = (-|x|^2/2) * 1   (first k-step of the seed block only)
              mma_ss<KIND>(t_s, kSdescSw128Hi | static_cast<uint64_t>(sd),
                           kSdescSw128Hi | static_cast<uint64_t>(sd + ((2 * TILE_BYTES) >> 4)), idesc_s, 0u);
            } else {
              step(0); step(1); step(2);
            }
            // one non-blocking probe: when the feed keeps up (the usual case) the next k-block starts
            // without a poll at its head; when it does not, this stage is still released first
            ready = more && mbar_test_wait(BAR(B_FULL + nstage), nphase) != 0;
            if (!no_smma && kb >= 0) step(3);
            tc_commit(BAR(B_EMPTY + stage));
            stage = nstage;
            phase = nphase;
          }
          tc_commit(BAR(B_SFULL + (n & 1)));
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
  } else if (warp >= 4 && !(p.dbg & 64)) {
    // ======================= epilogue =======================
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;          // row inside the block
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float nsl2 = p.neg_scale_log2;
    // operand scalings (powers of two written by the pre-pass; 1 for KIND_TF32)
    const float s_r = __ldg(p.r_scale), s_q = __ldg(p.q_scale);
    const float m2inv = -2.f / (s_r * s_q);        // acc -> -2 (x.c - rho |x|^2 / 2)
    const float one_m_rho = 1.f - s_r / s_q;       // 0 when both point sets share a scale
    uint32_t n = 0;
    uint32_t item_cnt = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++item_cnt) {
      const WorkItem w = decode_item(p, it);
      const int grow = w.row0 + row;
      const float rn = ((grow < p.n_rows) ? __ldg(p.rnorm + grow) : 0.f) * one_m_rho;
      for (int j = w.jt0; j < w.jt1; ++j, ++n) {
        const uint32_t b = n & 1;
        mbar_wait_warp(BAR(B_SFULL + b), (n >> 1) & 1);
        tc_fence_after();
        const uint32_t t_s = tmem_base + lane_off + (b ? TM_S1 : TM_S0);
        const uint32_t t_lo = tmem_base + lane_off + TM_PLO;
        const int col0 = j * BN;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t s[32];
          __syncwarp();
          tmem_ld32(t_s + ch * 32, s);
          float qn[32];
          const float4* qp = reinterpret_cast<const float4*>(p.qnorm + col0 + ch * 32);
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 t = __ldg(qp + v);
            qn[4 * v + 0] = t.x; qn[4 * v + 1] = t.y; qn[4 * v + 2] = t.z; qn[4 * v + 3] = t.w;
          }
          tc_wait_ld();
          if (p.dbg & 4) {
          } else if (mode_mmv) {
            uint32_t lo[32];
            uint32_t ph[SPILL16 ? 16 : 1], pl[SPILL16 ? 8 : 1];
            if (SPILL16) {
#pragma unroll
              for (int i = 0; i < 8; ++i) pl[i] = 0u;
            }
            // tf32 split on the integer pipe: hi = K rounded to 11 bits (add half an ulp, clear 13 bits), lo = K - hi
            // exactly; the tensor core reads only the upper 19 bits of lo (2^-21 K).  Two cvt.rna per element
            // would share the 16-lane XU pipe with ex2 and made the epilogue the bottleneck at d <= 256.
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              float d0 = fmaf(m2inv, __uint_as_float(s[c]), rn + qn[c]);
              float d1 = fmaf(m2inv, __uint_as_float(s[c + 1]), rn + qn[c + 1]);
              d0 = fmaxf(d0, 0.f);
              d1 = fmaxf(d1, 0.f);
              const float k0 = ex2_approx(d0 * nsl2), k1 = ex2_approx(d1 * nsl2);
              const uint32_t h0 = (__float_as_uint(k0) + 0x1000u) & 0xFFFFE000u;
              const uint32_t h1 = (__float_as_uint(k1) + 0x1000u) & 0xFFFFE000u;
              s[c] = h0;
              s[c + 1] = h1;
              lo[c] = __float_as_uint(k0 - __uint_as_float(h0));
              lo[c + 1] = __float_as_uint(k1 - __uint_as_float(h1));
              if (SPILL16) {
                const __half2 h = __floats2half2_rn(k0, k1);
                const float2 hf = __half22float2(h);
                ph[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                pl[c >> 2] |= (lo8_of(k0 - hf.x) << (8 * (c & 3))) | (lo8_of(k1 - hf.y) << (8 * (c & 3) + 8));
              }
            }
            if (ch == 0 && n > 0) mbar_wait_warp(BAR(B_PVDONE), (n - 1) & 1);  // K_lo buffer free
            tmem_st32(t_s + ch * 32, s);
            tmem_st32(t_lo + ch * 32, lo);
            if (SPILL16) {
              // [plane][column tile][row block][group of 8 centres][128 rows][8]: a warp store covers 32 rows x 16 B =
              // 512 contiguous bytes (with one row per thread and row-major tiles every STG touched 32 lines and the
              // LSU, not HBM, bounded the spill).  Rows past n_rows of the last block come from zero-filled
              // operands (finite K) and are written too; the matching W16 rows are zero.
              const int64_t e0 = (((static_cast<int64_t>(j) * p.n_rowblocks + (w.row0 >> 7)) * 16 + ch * 4) * 128 + row) * 8;
              __half* dst = p.panel16 + e0;
              uint8_t* dlo = reinterpret_cast<uint8_t*>(p.panel16 + p.panel16_plane) + e0;     // one byte per value
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                *reinterpret_cast<uint4*>(dst + v * 1024) = make_uint4(ph[4 * v], ph[4 * v + 1], ph[4 * v + 2], ph[4 * v + 3]);
                *reinterpret_cast<uint2*>(dlo + v * 1024) = make_uint2(pl[2 * v], pl[2 * v + 1]);
              }
            }
            if (!SPILL16 && p.panel != nullptr && grow < p.n_rows) {
              // spill K = K_hi + K_lo (exactly what the contraction above uses) to the row panel
              float* prow = p.panel + static_cast<int64_t>(grow) * p.ldpanel + col0 + ch * 32;
#pragma unroll
              for (int v = 0; v < 8; ++v) {
                float4 t;
                t.x = __uint_as_float(s[4 * v + 0]) + __uint_as_float(lo[4 * v + 0]);
                t.y = __uint_as_float(s[4 * v + 1]) + __uint_as_float(lo[4 * v + 1]);
                t.z = __uint_as_float(s[4 * v + 2]) + __uint_as_float(lo[4 * v + 2]);
                t.w = __uint_as_float(s[4 * v + 3]) + __uint_as_float(lo[4 * v + 3]);
                *reinterpret_cast<float4*>(prow + 4 * v) = t;
              }
            }
          } else if (LINEAR) {
            // MODE_STORE, linear: out = alpha (x . c) + beta out
            const float a_ss = p.lin_alpha / (s_r * s_q);
            if (grow < p.n_rows) {
              float* orow = p.out + static_cast<int64_t>(grow) * p.ldo;
              const int c0 = col0 + ch * 32;
              if (p.store_vec4 && c0 + 32 <= p.n_cols) {
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                  float4 t;
                  t.x = a_ss * __uint_as_float(s[4 * v + 0]); t.y = a_ss * __uint_as_float(s[4 * v + 1]);
                  t.z = a_ss * __uint_as_float(s[4 * v + 2]); t.w = a_ss * __uint_as_float(s[4 * v + 3]);
                  float4* o4 = reinterpret_cast<float4*>(orow + c0 + 4 * v);
                  if (p.lin_beta != 0.f) {
                    const float4 o = *o4;
                    t.x = fmaf(p.lin_beta, o.x, t.x); t.y = fmaf(p.lin_beta, o.y, t.y);
                    t.z = fmaf(p.lin_beta, o.z, t.z); t.w = fmaf(p.lin_beta, o.w, t.w);
                  }
                  *o4 = t;
                }
              } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  if (c0 + c < p.n_cols) {
                    float v = a_ss * __uint_as_float(s[c]);
                    if (p.lin_beta != 0.f) v = fmaf(p.lin_beta, orow[c0 + c], v);
                    orow[c0 + c] = v;
                  }
                }
              }
            }
          } else {
            // MODE_STORE: write K straight to global (row-major, ld = ldo)
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              float d2 = fmaf(m2inv, __uint_as_float(s[c]), rn + qn[c]);
              d2 = fmaxf(d2, 0.f);
              s[c] = __float_as_uint(ex2_approx(d2 * nsl2));
            }
            if (grow < p.n_rows) {
              float* orow = p.out + static_cast<int64_t>(grow) * p.ldo;
              const int c0 = col0 + ch * 32;
              if (p.store_vec4 && c0 + 32 <= p.n_cols) {
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                  float4 t;
                  t.x = __uint_as_float(s[4 * v + 0]); t.y = __uint_as_float(s[4 * v + 1]);
                  t.z = __uint_as_float(s[4 * v + 2]); t.w = __uint_as_float(s[4 * v + 3]);
                  *reinterpret_cast<float4*>(orow + c0 + 4 * v) = t;
                }
              } else {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (c0 + c < p.n_cols) orow[c0 + c] = __uint_as_float(s[c]);
              }
            }
          }
        }
        if (mode_mmv) tc_wait_st();
        tc_fence_before();
        mbar_arrive(PREADY(b));
      }
      if (mode_mmv) {
        // W for this item is complete once the last tile's contraction has retired.
        mbar_wait_warp(BAR(B_WFULL), item_cnt & 1);
        tc_fence_after();
        const uint32_t t_w = tmem_base + lane_off + TM_W;
        float* orow = p.out + static_cast<int64_t>(w.split) * p.split_stride +
                      static_cast<int64_t>(grow) * T_pad;
        for (int c0 = 0; c0 < T_pad; c0 += 16) {
          uint32_t r[16];
          __syncwarp();
          tmem_ld16(t_w + c0, r);
          tc_wait_ld();
          if (grow < p.n_rows) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              float4 t;
              t.x = __uint_as_float(r[4 * v + 0]); t.y = __uint_as_float(r[4 * v + 1]);
              t.z = __uint_as_float(r[4 * v + 2]); t.w = __uint_as_float(r[4 * v + 3]);
              *reinterpret_cast<float4*>(orow + c0 + 4 * v) = t;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(BAR(B_WEMPTY));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TM_COLS);
  if ((p.dbg & 16) && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    const long long c1 = clock64();
    printf("tile dbg[%d]: cta %d cycles=%lld ns=%llu  => %.1f MHz\n", p.dbg, blockIdx.x, c1 - dbg_c0, t1 - dbg_t0,
           1e3 * double(c1 - dbg_c0) / double(t1 - dbg_t0));
  }
}

// ----------------------------------------------------------------------------- host side
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// 2-D tensor [rows x cols] of fp32 (esize 4) or fp16 (esize 2) with row pitch ld (elements);
// box = [box_rows x 128 bytes], 128B swizzle.
int make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld,
             int box_rows, int esize = 4) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(ODF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esize};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esize), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld",
             static_cast<int>(r), (long long)rows, (long long)cols, (long long)ld);
    return set_error(ODF_ERR_CUDA, buf);
  }
  return ODF_OK;
}

}  // namespace

// [rows x cols] of fp16 / fp32 with row pitch ld (elements); box = [box_rows x 128 bytes], SWIZZLE_128B.
int make_map_sw128(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int esize) {
  return make_map(m, base, rows, cols, ld, box_rows, esize);
}

// Un-swizzled 2-D fp16 map: [rows x cols] with pitch ld (elements), box = [box_rows x box_cols].
int make_map_plain_f16(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(ODF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(ODF_ERR_CUDA, "cuTensorMapEncodeTiled (plain fp16 map) failed");
  return ODF_OK;
}

// The same for bytes (the lo plane of the K panel).
int make_map_plain_u8(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(ODF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(ODF_ERR_CUDA, "cuTensorMapEncodeTiled (plain byte map) failed");
  return ODF_OK;
}

// Plain (un-swizzled) 2-D fp32 map: box = [box_rows x box_cols]; out-of-bounds elements read as 0.
int make_map_plain_f32(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                       int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(ODF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(ODF_ERR_CUDA, "cuTensorMapEncodeTiled (plain fp32 map) failed");
  return ODF_OK;
}

namespace {

int num_sms() { return device_sm_count(); }

}  // namespace

// `row_bytes` = bytes of one point in ONE of the hi / lo arrays
int tile_default_splits(int64_t n_rows, int64_t n_cols, int64_t row_bytes) {
  const int sms = num_sms();
  const int64_t n_rb = (n_rows + BM - 1) / BM;
  const int64_t n_ct = (n_cols + BN - 1) / BN;
  int64_t splits = 1;
  // (a) not enough row blocks to fill the machine a few times over: split the column range
  if (n_rb < 8 * sms) splits = (8 * sms + n_rb - 1) / n_rb;
  // (b) row blocks too fat for L2 when every SM streams its own: make ~12 CTAs share one
  const int64_t rb_bytes = static_cast<int64_t>(BM) * row_bytes * 2;
  if (rb_bytes * sms > (48ll << 20) && splits < 12) splits = 12;
  // (c) the W accumulator lives in TMEM across the column tiles of one item and the tensor core adds with
  // truncation: the bias grows with the chain length (measured 1.3e-4 relative at 235 tiles, 5e-6 at 7),
  // so a chain is at most 8 tiles; the partial slabs are summed in fp32 round-to-nearest by odf_finish_*
  if (splits * 8 < n_ct) splits = (n_ct + 7) / 8;
  // keep at least 4 column tiles per item so the pipeline fill/drain stays amortised
  const int64_t max_splits = n_ct >= 4 ? n_ct / 4 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  // normalise so that no split is empty (the launcher insists on it)
  const int64_t tps = (n_ct + splits - 1) / splits;
  splits = (n_ct + tps - 1) / tps;
  return static_cast<int>(splits);
}

int launch_gauss_tile(const TileLaunch& L, cudaStream_t stream) {
  if (L.kind != KIND_TF32 && L.kind != KIND_F16) return set_error(ODF_ERR_ARG, "unknown operand kind");
  if (tile2_eligible(L)) return launch_gauss_tile2(L, stream);
  const int64_t BK = kblock_elems(L.kind);
  const int esize = L.kind == KIND_F16 ? 2 : 4;
  const int64_t pitch = L.d_pad + 2 * BK;
  if (L.d_pad % BK != 0 || L.d_pad <= 0) return set_error(ODF_ERR_ARG, "d_pad must be a positive multiple of the k-block width");
  if (L.n_rows <= 0 || L.n_cols <= 0) return set_error(ODF_ERR_ARG, "empty operand");
  if (L.mode == MODE_MMV && !(L.T_pad == 16 || L.T_pad == 32))
    return set_error(ODF_ERR_ARG, "T_pad must be 16 or 32");
  static DeviceOnce attr_once;
  bool& attr_set = attr_once.here();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gauss_tile_kernel<KIND_TF32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gauss_tile_kernel<KIND_F16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gauss_tile_kernel<KIND_TF32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gauss_tile_kernel<KIND_F16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gauss_tile_kernel<KIND_TF32, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gauss_tile_kernel<KIND_F16, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gauss_tile_kernel)");
    attr_set = true;
  }
  if (L.linear && (L.mode != MODE_STORE || L.panel != nullptr || L.panel16 != nullptr))
    return set_error(ODF_ERR_ARG, "the linear store variant is MODE_STORE without a spill");
  CUtensorMap mRh, mRl, mQh, mQl, mVh, mVl;
  int rc;
  if ((rc = make_map(&mRh, L.r_hi, L.n_rows, pitch, pitch, BM, esize))) return rc;
  if ((rc = make_map(&mRl, L.r_lo, L.n_rows, pitch, pitch, BM, esize))) return rc;
  if ((rc = make_map(&mQh, L.q_hi, L.n_cols, pitch, pitch, BN, esize))) return rc;
  if ((rc = make_map(&mQl, L.q_lo, L.n_cols, pitch, pitch, BN, esize))) return rc;
  if (L.mode == MODE_MMV) {
    if ((rc = make_map(&mVh, L.vt_hi, L.T_pad, L.ldvt, L.ldvt, L.T_pad))) return rc;
    if ((rc = make_map(&mVl, L.vt_lo, L.T_pad, L.ldvt, L.ldvt, L.T_pad))) return rc;
  } else {
    mVh = mRh;
    mVl = mRl;
  }
  TileParams p;
  p.n_rows = static_cast<int>(L.n_rows);
  p.n_cols = static_cast<int>(L.n_cols);
  p.kblocks = static_cast<int>(L.d_pad / BK);
  p.T_pad = L.T_pad;
  p.mode = L.mode;
  p.n_rowblocks = static_cast<int>((L.n_rows + BM - 1) / BM);
  p.n_coltiles = static_cast<int>((L.n_cols + BN - 1) / BN);
  int splits = L.n_splits > 0 ? L.n_splits : 1;
  if (splits > p.n_coltiles) splits = p.n_coltiles;
  p.tiles_per_split = (p.n_coltiles + splits - 1) / splits;
  p.n_splits = (p.n_coltiles + p.tiles_per_split - 1) / p.tiles_per_split;
  if (L.mode == MODE_MMV && p.n_splits != L.n_splits)
    return set_error(ODF_ERR_ARG, "n_splits must divide the column tiles without empty splits (use odf_tile_splits)");
  const int sms = num_sms();
  p.group_rows = sms / p.n_splits;
  if (p.group_rows < 1) p.group_rows = 1;
  p.neg_scale_log2 = static_cast<float>(-1.4426950408889634 / (2.0 * double(L.sigma) * double(L.sigma)));
  p.rnorm = L.r_norm;
  p.qnorm = L.q_norm;
  p.r_scale = L.r_scale;
  p.q_scale = L.q_scale;
  p.out = L.out;
  p.ldo = L.ldo;
  p.split_stride = L.split_stride;
  p.panel = L.panel;
  p.ldpanel = L.ldpanel;
  if (L.panel != nullptr && (L.mode != MODE_MMV || L.ldpanel < static_cast<int64_t>(p.n_coltiles) * BN ||
                             L.ldpanel % 4 != 0 || (reinterpret_cast<uintptr_t>(L.panel) & 15) != 0))
    return set_error(ODF_ERR_ARG, "panel must be 16-byte aligned with pitch >= round_up(n_cols, 128)");
  p.panel16 = static_cast<__half*>(L.panel16);
  p.panel16_plane = static_cast<int64_t>(p.n_coltiles) * p.n_rowblocks * 16384;
  if (L.panel16 != nullptr && (L.mode != MODE_MMV || L.panel != nullptr || (reinterpret_cast<uintptr_t>(L.panel16) & 127) != 0))
    return set_error(ODF_ERR_ARG, "panel16 must be 128-byte aligned, MODE_MMV only, and excludes the fp32 panel");
  {
    const char* e = getenv("ODF_TILE_DEBUG");
    p.dbg = e ? atoi(e) : 0;
  }
  p.store_vec4 = (L.ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(L.out) & 15) == 0) ? 1 : 0;
  p.lin_alpha = L.lin_alpha;
  p.lin_beta = L.lin_beta;
  const int n_items = p.n_rowblocks * p.n_splits;
  const int grid = n_items < sms ? n_items : sms;
  if (L.linear && L.kind == KIND_F16)
    gauss_tile_kernel<KIND_F16, 0, 1><<<grid, 256, SMEM_BYTES, stream>>>(mRh, mRl, mQh, mQl, mVh, mVl, p);
  else if (L.linear)
    gauss_tile_kernel<KIND_TF32, 0, 1><<<grid, 256, SMEM_BYTES, stream>>>(mRh, mRl, mQh, mQl, mVh, mVl, p);
  else if (L.kind == KIND_F16 && L.panel16 != nullptr)
    gauss_tile_kernel<KIND_F16, 1><<<grid, 256, SMEM_BYTES, stream>>>(mRh, mRl, mQh, mQl, mVh, mVl, p);
  else if (L.kind == KIND_F16)
    gauss_tile_kernel<KIND_F16, 0><<<grid, 256, SMEM_BYTES, stream>>>(mRh, mRl, mQh, mQl, mVh, mVl, p);
  else if (L.panel16 != nullptr)
    gauss_tile_kernel<KIND_TF32, 1><<<grid, 256, SMEM_BYTES, stream>>>(mRh, mRl, mQh, mQl, mVh, mVl, p);
  else
    gauss_tile_kernel<KIND_TF32, 0><<<grid, 256, SMEM_BYTES, stream>>>(mRh, mRl, mQh, mQl, mVh, mVl, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "gauss_tile_kernel launch");
  return ODF_OK;
}

}  // namespace odf
