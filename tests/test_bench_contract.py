"""CPU tests of bench.py's contract: the reference arm's JSON line (oracle port timed on the host cores), rank > 0
staying silent, the GPU arm refusing to run without a device, and the arithmetic of the roofline entries on synthetic
CUDA-event records (no GPU needed)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], {"ODF_CPU_SAMPLE": "3000,200"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "falkon_fit_gflops" and j["unit"] == "GFLOP/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 1
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "3000" in j["cpu_baseline"]["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "N=3000" in j["config"]["workload"] and "model" not in j["config"]


def test_reference_arm_is_rank0_only():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], {"RANK": "1", "WORLD_SIZE": "2", "ODF_CPU_SAMPLE": "3000,200"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "1"], timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


class _Ev:
    def __init__(self, t):
        self.t = t

    def elapsed_time(self, other):
        return other.t - self.t


def test_roofline_entries_from_event_records():
    import bench
    peaks = {"hbm_gbs": 6500.0, "bf16_tflops_sustained": 1400.0}
    traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
    # 4 panel launches of a 524288 x 10000 panel at 3 ms (2 per kernel) and 2 of the ragged 475712-row one
    pe = []
    t = 0.0
    for n, name in ((524288, "panel16_kernel"), (524288, "panel16_mmv_kernel"), (475712, "panel16_kernel"),
                    (475712, "panel16_mmv_kernel"), (524288, "panel16_kernel"), (524288, "panel16_mmv_kernel")):
        pe.append((_Ev(t), _Ev(t + 3.0), n, 10000, 32, name))
        t += 3.0
    r = bench.panel_roofline(pe, steps=1, step_ms=36.0, peaks=peaks, traffic_json=traffic)
    alg = 3.0 * 10112 * (4 * 524288 + 2 * 475776)
    assert r["bound"] == "hbm" and r["kernel"] == "panel16_kernel + panel16_mmv_kernel" and r["launches_timed"] == 6
    assert abs(r["achieved"] - alg / 18.0 / 1e6) < 1e-6 * r["achieved"] and abs(r["frac"] - r["achieved"] / 6500.0) < 1e-12
    assert abs(r["share_of_step"] - 0.5) < 1e-12 and abs(r["algorithmic_bytes_per_launch"] - alg / 6) < 1
    # traffic: the captured shape launched most often (524288 rows), read + write of that kernel's capture
    assert r["traffic"] in (15971971000 + 23584000, 15906150000 + 24156000) and "524288 x 10000" in r["traffic_source"]
    assert 0.99 < r["traffic"] / (3.0 * 10112 * 524288) < 1.01
    assert set(r["per_kernel"]) == {"panel16_kernel", "panel16_mmv_kernel"} and r["per_kernel"]["panel16_kernel"]["launches"] == 3
    # tile: 2 launches of 524288 x 10000 x 1024, T = 30, 24 ms each
    te = [(_Ev(0.0), _Ev(24.0), 524288, 10000, 1024, 30), (_Ev(24.0), _Ev(48.0), 524288, 10000, 1024, 30)]
    q = bench.tile_roofline(te, steps=1, step_ms=480.0, peaks=peaks, traffic_json=traffic)
    flops = 2.0 * 524288 * 10000 * (1024 + 30)
    assert q["bound"] == "tensor" and abs(q["achieved"] - flops / 24.0 / 1e9) < 1e-6 * q["achieved"]
    assert abs(q["executed_tensor_tflops"] - 6.0 * 524288 * 10000 * (1024 + 32) / 24.0 / 1e9) < 1e-6 * q["executed_tensor_tflops"]
    assert abs(q["share_of_step"] - 0.1) < 1e-12 and q["traffic"] == 24885498000 + 16680812000
    # a shape without a capture reports no traffic
    q2 = bench.tile_roofline([(_Ev(0.0), _Ev(1.0), 4096, 1000, 256, 21)], 1, 10.0, peaks, traffic)
    assert q2["traffic"] is None and q2["traffic_source"] is None
    # no MEASURED_PEAKS.json: the stated fallbacks
    assert bench.panel_roofline(pe, 1, 36.0, {}, {})["peak"] == 6650.0 and bench.tile_roofline(te, 1, 480.0, {}, {})["peak"] == 1400.0


def test_fit_flops_matches_the_survey_formula():
    import bench
    N, M, d, T = 1_000_000, 10_000, 1024, 30
    f_mmv = 2.0 * N * M * d + 2.0 * N * M * T
    f_dmmv = 2.0 * N * M * d + 4.0 * N * M * T
    assert bench.fit_flops(N, M, d, T) == f_mmv + 22 * f_dmmv
    assert abs(f_dmmv - 2.168e13) < 1e10                           # C2 F_dmmv = 2.168e13
