"""Time one resident sweep (two-pass vs one-pass) on a C2-shaped chunk set; prints ms per sweep and panel GB/s.

    python tools/sweep_time.py [N] [M] [d] [T]          ODF_SWEEP_LAG / ODF_SWEEP_POLICY_A / ODF_SWEEP_POLICY_C tune the one-pass kernel
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "online-detection_b200"))
import odf  # noqa: E402
from odf import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
T = int(sys.argv[4]) if len(sys.argv) > 4 else 30
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(N, d, device="cuda", generator=g) * 0.7
C = X[torch.randperm(N, device="cuda", generator=g)[:M]].contiguous()
sigma = 20.0 * (d / 1024) ** 0.5
k = odf.GaussianKernel(sigma)
cols = k._prep(C)
rows = k._prep(X, like=cols)
del X
y = torch.randn(N, T, device="cuda", generator=g)
res = {}
for fused in ([False, True] if os.environ.get("ODF_SWEEP_ONLY") is None else [True]):
    ops.RESIDENT_FUSED = fused
    sw = ops.Sweeper(rows, cols, sigma, T, mode="resident")
    out = torch.empty((M, T), device="cuda")
    sw.dmmv(None, y, out, 1.0, 1.0 / N)
    v = torch.randn(M, T, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    for _ in range(2):
        sw.dmmv(v, None, out, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        sw.dmmv(v, None, out, 1.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    panel_gb = sum(int(ops._lib.load().odf_panel16_bytes(r1 - r0, M)) for (r0, r1) in sw.chunks) / 1e9
    res[fused] = out.clone()
    print("%s sweep: %.3f ms   (panel %.1f GB: %.0f GB/s per single read)" % ("one-pass" if fused else "two-pass", ms, panel_gb, panel_gb / ms * 1e3), flush=True)
    del sw
    torch.cuda.empty_cache()
if len(res) == 2:
    a, b = res[False].double(), res[True].double()
    print("one-pass vs two-pass: max rel col err %.2e" % float(((a - b).abs().max(0).values / a.abs().max(0).values).max()))
