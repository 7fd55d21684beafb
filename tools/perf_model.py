#!/usr/bin/env python
"""First-principles time model of the resident FALKON fit, to be read next to the measured points.
Not a measurement: it multiplies the algorithmic bytes / flops of each phase by the rates MEASURED for the kernels on
one B200 (profiles/r1_bench_v12_c2.json, r1_ncu_resident_v12_summary.md, r1_bench_v6_*), and prints what that gives at
1 / 2 / 4 / 8 GPUs so that the driver's round-end scaling run can be checked against it.

    python tools/perf_model.py [--N 1000000 --M 10000 --d 1024 --T 30]
"""
import argparse

# measured rates (one B200, power-capped clocks as in the bench runs)
PANEL_GBS = 6600.0            # panel16_kernel / panel16_mmv_kernel inside a fit (CUDA events), GB/s
TILE_ALG_TFLOPS = 369.0       # fused tile with spill, algorithmic TFLOP/s at d = 1024 (2 r c (d + T) per launch)
SMALL_PER_SWEEP_MS = 0.9      # finish_w16 x 2, finish_rows, V16 conversion, per sweep (C2, one GPU), scales with rows
APPLY_PER_ITER_MS = 0.75      # four applications of the explicit inverses (replicated), per CG iteration at M = 10 k
VEC_PER_ITER_MS = 0.12        # CG vector kernels per iteration
PRECOND_MS = {1: 80.0, 2: 74.0, 4: 69.0, 8: 64.0}     # measured builds at M = 10 k (replicated below 4 ranks, column blocks from 4)
ALLREDUCE_MS = 0.05           # M x T fp32 over NVLink, per sweep
MEASURED_FIT_S = {1: 0.441, 2: 0.270}                 # resident fit, C2 (profiles/r1_bench_v15_c2.json, r1_bench_v9_2gpu.json)
MEASURED_STREAMED_S = {1: 1.52, 2: 0.82, 4: 0.43, 8: 0.265}


def model(N, M, d, T, world):
    rows = N / world
    m_pad = (M + 127) // 128 * 128
    panel_bytes = 4.0 * rows * m_pad
    sweeps = 23
    fill_ms = 2.0 * rows * M * (d + T) / (TILE_ALG_TFLOPS * 1e12) * 1e3
    # sweep 0: tile + one panel pass; sweeps 1..22: two panel passes
    panel_ms = (1 + 2 * (sweeps - 1)) * panel_bytes / (PANEL_GBS * 1e9) * 1e3
    scale_m = (M / 10000.0) ** 2
    small_ms = sweeps * SMALL_PER_SWEEP_MS * (rows / 1e6)
    apply_ms = 22 * APPLY_PER_ITER_MS * scale_m / (world if world > 1 else 1) + (22 * 0.06 if world > 1 else 0.0)
    vec_ms = 20 * VEC_PER_ITER_MS
    comm_ms = sweeps * ALLREDUCE_MS if world > 1 else 0.0
    precond_ms = PRECOND_MS.get(world, PRECOND_MS[8]) * (M / 10000.0) ** 3
    prepare_ms = 4.6 * (rows / 1e6) * (d / 1024.0) + 0.3
    cg_ms = fill_ms + panel_ms + small_ms + apply_ms + vec_ms + comm_ms
    return {"prepare": prepare_ms, "precond": precond_ms, "fill (tile)": fill_ms, "panel passes": panel_ms,
            "small + applies + vec + comm": small_ms + apply_ms + vec_ms + comm_ms, "fit": prepare_ms + precond_ms + cg_ms}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1_000_000)
    ap.add_argument("--M", type=int, default=10_000)
    ap.add_argument("--d", type=int, default=1024)
    ap.add_argument("--T", type=int, default=30)
    a = ap.parse_args()
    print("resident FALKON fit, N=%d M=%d d=%d T=%d: model from measured kernel rates (ms)" % (a.N, a.M, a.d, a.T))
    print("| GPUs | prepare | precond | fill (tile) | panel passes | small+applies+vec+comm | fit (model) | fit (measured) | streamed fit (measured) | speed-up vs 1 GPU (model) |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    base = None
    for w in (1, 2, 4, 8):
        r = model(a.N, a.M, a.d, a.T, w)
        base = base or r["fit"]
        c2 = (a.N, a.M, a.d, a.T) == (1_000_000, 10_000, 1024, 30)
        meas = ("%.0f" % (MEASURED_FIT_S[w] * 1e3)) if (c2 and w in MEASURED_FIT_S) else "-"
        stream = ("%.0f" % (MEASURED_STREAMED_S[w] * 1e3)) if c2 else "-"
        print("| %d | %.1f | %.1f | %.1f | %.1f | %.1f | **%.0f** | %s | %s | %.2fx |"
              % (w, r["prepare"], r["precond"], r["fill (tile)"], r["panel passes"], r["small + applies + vec + comm"], r["fit"],
                 meas, stream, base / r["fit"]))


if __name__ == "__main__":
    main()
