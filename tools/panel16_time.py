"""Time the two panel kernels alone on a zero-filled panel of a C2-sized chunk (the content does not matter for the timing).

    python tools/panel16_time.py [rows] [M]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "online-detection_b200"))
from odf import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
M = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
L = ops._lib.load()
dev = torch.device("cuda")
nbytes = int(L.odf_panel16_bytes(n, M))
p16 = torch.zeros((nbytes,), dtype=torch.uint8, device=dev)
W16 = torch.zeros(((n + 127) // 128 * 128, 64), dtype=torch.float16, device=dev)
V16 = torch.zeros(((M + 127) // 128 * 128, 64), dtype=torch.float16, device=dev)
absmax = torch.zeros(32, dtype=torch.int32, device=dev)
S = int(L.odf_panel16_splits(n, M))
out = torch.empty((S, M, 32), device=dev)
Sv = int(L.odf_panel16_mmv_splits(n, M))
outv = torch.empty((Sv, n, 32), device=dev)
for name, fn in (("panel16_kernel (K^T w)", lambda: ops.panel16_tmm(p16, W16, absmax, n, M, out)),
                 ("panel16_mmv_kernel (K v)", lambda: ops.panel16_mmv(p16, V16, absmax, n, M, outv)),
                 ("panel16_kernel<hi only>", lambda: ops.panel16_tmm(p16, W16, absmax, n, M, out, hi_only=True))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    b = nbytes * (2.0 / 3.0 if "hi only" in name else 1.0)
    print("%-28s %d x %d: %.3f ms  %.0f GB/s of panel" % (name, n, M, ms, b / ms / 1e6), flush=True)
