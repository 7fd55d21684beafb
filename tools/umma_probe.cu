// Micro-probe for tcgen05.mma issue/throughput on sm_100a (bring-up tool, not part of libodf).
// One CTA per SM; an elected lane of warp 0 issues a stream of MMAs over resident (garbage)
// shared-memory operands and the CTA reports cycles per MMA for a set of shapes:
//   N (64..256), same vs alternating accumulators, A from shared memory (SS) vs tensor memory (TS),
//   kind::f16 vs kind::tf32.  The numbers calibrate the tile design in DESIGN.md §3.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../online-detection_b200/csrc/odf_ptx.cuh"

using namespace odf;

constexpr int TILE = 128 * 128;          // 16 KB operand tile (128 rows x 128 B)
constexpr int SMEM = 13 * TILE + 1024 + 64;

// PATTERN 0: all MMAs into one accumulator;  1: alternate between two;  2: three
template <int KIND, int N, int TS, int ALT>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 13 * TILE);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 12 * TILE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    const uint32_t idesc = KIND == 1 ? make_idesc_f16(128, N) : make_idesc_tf32(128, N);
    const uint32_t sd = (smem_u32(smem) & 0x3FFFFu) >> 4;
    long long c0 = 0, c1 = 0;
    if (elect_one()) {
      c0 = clock64();
      for (int it = 0; it < iters; ++it) {
        // walk 3 "stages" of (A_hi, A_lo, B_hi, B_lo)-like tiles the way the real kernel does
        const uint32_t st = sd + (it % 3) * ((4 * TILE) >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            const uint64_t a = kSdescSw128Hi | static_cast<uint64_t>(st + (((t == 0 ? 1 : 0) * TILE + ks * 32) >> 4));
            const uint64_t b = kSdescSw128Hi | static_cast<uint64_t>(st + (((t == 1 ? 3 : 2) * TILE + ks * 32) >> 4));
            const int sel = ALT == 0 ? 0 : (ALT == 1 ? ((ks * 3 + t) & 1) : t);
            const uint32_t d = tmem + (N <= 128 ? sel * 128 : (sel & 1) * 256);
            if (TS) {
              const uint32_t at = tmem + 384 + ((ks * 3 + t) % 4) * 8;
              if (KIND == 1) mma_f16_ts(d, at, b, idesc, 1u);
              else mma_tf32_ts(d, at, b, idesc, 1u);
            } else {
              mma_ss<KIND>(d, a, b, idesc, 1u);
            }
          }
        }
      }
      tc_commit(smem_u32(bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(bar), 0);
    if (elect_one()) {
      c1 = clock64();
      out[blockIdx.x] = c1 - c0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int KIND, int N, int TS, int ALT>
void run(const char* name, int iters, long long* dout) {
  auto k = probe<KIND, N, TS, ALT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  k<<<sms, 128, SMEM>>>(iters, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  k<<<sms, 128, SMEM>>>(iters, dout);
  cudaDeviceSynchronize();
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), dout, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0; long long mx = 0;
  for (long long v : h) { s += double(v); if (v > mx) mx = v; }
  const double per = s / sms / (12.0 * iters);
  printf("%-44s cycles/MMA avg %.1f (max CTA %.1f)  floor %.0f  => %.0f%% of floor rate\n", name, per,
         double(mx) / (12.0 * iters), N / 2.0, 100.0 * (N / 2.0) / per);
}


// Variant closer to the real tile: per k-block (12 MMAs) the issuer waits on a "full" barrier that a
// producer warp arrives on (HS), commits to an "empty" barrier, and (EPI) four epilogue warps stream
// tcgen05.ld / tcgen05.st over the other TMEM columns at the real kernel's per-tile volume.
template <int KIND, int HS, int EPI, int PIPE>
__global__ void __launch_bounds__(256, 1) probe2(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 13 * TILE);   // [0..2] full, [3..5] empty, [6] done
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  volatile int* stopflag = reinterpret_cast<volatile int*>(slot + 1);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 12 * TILE / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (iters < 0) {   // negative iters: random fp16 payload (|x| < 2) instead of zeros
      uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      v = (h & 0x83FF83FFu) | 0x38003800u;
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (iters < 0) iters = -iters;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 7; ++i) mbar_init(smem_u32(bars + i), 1);
    *stopflag = 0;
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  auto BAR = [&](int i) { return smem_u32(bars + i); };
  if (warp == 0 && HS) {
    // producer: waits empty, arrives full (no data movement)
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(BAR(3 + stage), phase ^ 1);
      if (elect_one()) mbar_arrive(BAR(stage));
      __syncwarp();
      if (++stage == 3) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = KIND == 1 ? make_idesc_f16(128, 128) : make_idesc_tf32(128, 128);
    const uint32_t sd = (smem_u32(smem) & 0x3FFFFu) >> 4;
    long long c0 = clock64();
    int stage = 0; uint32_t phase = 0;
    auto issue = [&](uint32_t st, int ks) {
      const uint64_t a_hi = kSdescSw128Hi | static_cast<uint64_t>(st + ((0 * TILE + ks * 32) >> 4));
      const uint64_t a_lo = kSdescSw128Hi | static_cast<uint64_t>(st + ((1 * TILE + ks * 32) >> 4));
      const uint64_t b_hi = kSdescSw128Hi | static_cast<uint64_t>(st + ((2 * TILE + ks * 32) >> 4));
      const uint64_t b_lo = kSdescSw128Hi | static_cast<uint64_t>(st + ((3 * TILE + ks * 32) >> 4));
      mma_ss<KIND>(tmem, a_lo, b_hi, idesc, 1u);
      mma_ss<KIND>(tmem, a_hi, b_lo, idesc, 1u);
      mma_ss<KIND>(tmem, a_hi, b_hi, idesc, 1u);
    };
    if (PIPE == 0) {
      for (int it = 0; it < iters; ++it) {
        if (HS) { mbar_wait(BAR(stage), phase); tc_fence_after(); }
        const uint32_t st = sd + stage * ((4 * TILE) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) issue(st, ks);
          if (HS) tc_commit(BAR(3 + stage));
          if (it == iters - 1) tc_commit(BAR(6));
        }
        __syncwarp();
        if (++stage == 3) { stage = 0; phase ^= 1; }
      }
    } else if (PIPE == 1) {
      // software-pipelined: the wait for the NEXT stage sits between the MMAs of the current one
      mbar_wait(BAR(0), 0);
      tc_fence_after();
      for (int it = 0; it < iters; ++it) {
        const uint32_t st = sd + stage * ((4 * TILE) >> 4);
        int nstage = stage + 1; uint32_t nphase = phase;
        if (nstage == 3) { nstage = 0; nphase ^= 1; }
        if (elect_one()) {
          issue(st, 0); issue(st, 1);
        }
        __syncwarp();
        if (it + 1 < iters) { mbar_wait(BAR(nstage), nphase); tc_fence_after(); }
        if (elect_one()) {
          issue(st, 2); issue(st, 3);
          tc_commit(BAR(3 + stage));
          if (it == iters - 1) tc_commit(BAR(6));
        }
        __syncwarp();
        stage = nstage; phase = nphase;
      }
    } else {
      // single elected lane runs the whole loop (no per-block elect / syncwarp); waits by that lane only
      if (elect_one()) {
        for (int it = 0; it < iters; ++it) {
          mbar_wait(BAR(stage), phase);
          tc_fence_after();
          const uint32_t st = sd + stage * ((4 * TILE) >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) issue(st, ks);
          tc_commit(BAR(3 + stage));
          if (it == iters - 1) tc_commit(BAR(6));
          if (++stage == 3) { stage = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    }
    mbar_wait(BAR(6), 0);
    const long long c1 = clock64();
    if (lane == 0) { out[blockIdx.x] = c1 - c0; *stopflag = 1; }
  } else if (warp >= 4 && EPI) {
    // epilogue-like TMEM traffic on columns [128, 384): per "tile" 4 x (ld32 S, st32 K_hi, st32 K_lo)
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t r[32];
    while (*stopflag == 0) {
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        tmem_ld32(tmem + lane_off + 128 + ch * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(ex2_approx(__uint_as_float(r[c]) * 0.5f));
        tmem_st32(tmem + lane_off + 128 + ch * 32, r);
        tmem_st32(tmem + lane_off + 256 + ch * 32, r);
      }
      tc_wait_st();
      if (EPI == 2) __nanosleep(4000);   // roughly the real duty cycle: one tile's epilogue per ~12k cycles
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

template <int KIND, int HS, int EPI, int PIPE>
void run2(const char* name, int iters, long long* dout) {
  auto k = probe2<KIND, HS, EPI, PIPE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int rep = 0; rep < 2; ++rep) {
    k<<<sms, 256, SMEM>>>(iters, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), dout, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0;
  for (long long v : h) s += double(v);
  printf("%-44s cycles/MMA avg %.1f\n", name, s / sms / (12.0 * (iters < 0 ? -iters : iters)));
}

// Latency probe: k MMAs + commit, wait for the mbarrier; and a bare commit (no MMAs).
template <int K>
__global__ void __launch_bounds__(128, 1) probe_lat(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 13 * TILE);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 12 * TILE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); mbar_init(smem_u32(bar + 1), 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(smem_u32(slot), 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc_f16(128, 128);
    const uint32_t sd = (smem_u32(smem) & 0x3FFFFu) >> 4;
    long long tot = 0;
    const int reps = 64;
    for (int r = 0; r < reps; ++r) {
      const long long c0 = clock64();
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const uint64_t a = kSdescSw128Hi | static_cast<uint64_t>(sd + ((i % 4) * 32 >> 4));
        const uint64_t b = kSdescSw128Hi | static_cast<uint64_t>(sd + ((2 * TILE + (i % 4) * 32) >> 4));
        mma_ss<1>(tmem, a, b, idesc, 1u);
      }
      tc_commit(smem_u32(bar));
      mbar_wait(smem_u32(bar), r & 1);
      tot += clock64() - c0;
    }
    out[blockIdx.x] = tot / reps;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// Barrier ping-pong between two warps (plain mbarrier arrive / try_wait): round-trip cycles.
__global__ void __launch_bounds__(128, 1) probe_pingpong(long long* out) {
  __shared__ uint64_t bars[2];
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); fence_barrier_init(); }
  __syncthreads();
  const int reps = 256;
  if (warp == 0 && elect_one()) {
    const long long c0 = clock64();
    for (int r = 0; r < reps; ++r) {
      mbar_arrive(smem_u32(&bars[0]));
      mbar_wait(smem_u32(&bars[1]), r & 1);
    }
    out[blockIdx.x] = (clock64() - c0) / reps;
  } else if (warp == 1 && elect_one()) {
    for (int r = 0; r < reps; ++r) {
      mbar_wait(smem_u32(&bars[0]), r & 1);
      mbar_arrive(smem_u32(&bars[1]));
    }
  }
}

template <int K>
void run_lat(long long* dout) {
  auto k = probe_lat<K>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  k<<<148, 128, SMEM>>>(dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("lat: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, dout, sizeof h, cudaMemcpyDeviceToHost);
  double s = 0; for (long long v : h) s += double(v);
  printf("latency: %2d MMAs (128x128x16 f16) + commit + wait = %.0f cycles\n", K, s / 148);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 1024 * sizeof(long long));
  const int iters = 4000;
#define RUN(K, N, TS, ALT) run<K, N, TS, ALT>("kind=" #K " N=" #N " TS=" #TS " ALT=" #ALT, iters, dout)
  RUN(1, 64, 0, 0); RUN(1, 128, 0, 0); RUN(1, 192, 0, 0); RUN(1, 256, 0, 0);
  RUN(1, 64, 0, 1); RUN(1, 128, 0, 1); RUN(1, 128, 0, 2); RUN(1, 256, 0, 1);
  RUN(1, 64, 1, 0); RUN(1, 128, 1, 0); RUN(1, 256, 1, 0); RUN(1, 128, 1, 1);
  RUN(0, 64, 0, 0); RUN(0, 128, 0, 0); RUN(0, 192, 0, 0); RUN(0, 256, 0, 0);
  RUN(0, 128, 0, 1); RUN(0, 128, 1, 0); RUN(0, 32, 1, 0); RUN(0, 32, 1, 1); RUN(0, 64, 1, 1);
#define RUN2(K, HS, EPI, PIPE) run2<K, HS, EPI, PIPE>("probe2 kind=" #K " HS=" #HS " EPI=" #EPI " PIPE=" #PIPE, iters, dout)
  RUN2(1, 0, 0, 0); RUN2(1, 1, 0, 0); RUN2(1, 1, 0, 1); RUN2(1, 1, 0, 2); RUN2(1, 1, 2, 0); RUN2(1, 1, 2, 1); RUN2(1, 1, 2, 2);
  run2<1, 1, 2, 2>("probe2 kind=1 HS=1 EPI=2 PIPE=2 RANDOM DATA", -20000, dout);
  run2<1, 0, 0, 2>("probe2 kind=1 HS=0 EPI=0 PIPE=2 RANDOM DATA", -20000, dout);
  run2<1, 1, 2, 2>("probe2 kind=1 HS=1 EPI=2 PIPE=2 zeros, long", 20000, dout);
  run2<0, 1, 2, 2>("probe2 kind=0 HS=1 EPI=2 PIPE=2 RANDOM DATA", -20000, dout);
  run_lat<0>(dout); run_lat<1>(dout); run_lat<2>(dout); run_lat<4>(dout); run_lat<12>(dout); run_lat<24>(dout);
  probe_pingpong<<<148, 128>>>(dout);
  cudaDeviceSynchronize();
  { long long h[148]; cudaMemcpy(h, dout, sizeof h, cudaMemcpyDeviceToHost); double t = 0; for (long long v : h) t += double(v);
    printf("mbarrier ping-pong round trip (arrive -> try_wait wake -> arrive -> try_wait wake) = %.0f cycles\n", t / 148); }
  return 0;
}
