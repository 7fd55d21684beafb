"""Time odf_precond_apply (csrc/odf_tri.cu) against the fp64 product; ODF_PRECOND_APPLY=cublas times round 1's sgemm."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "online-detection_b200"))
from odf import ops  # noqa: E402

for M, T in ((10000, 30), (10000, 32), (4000, 30), (2000, 1), (30000, 21)):
    g = torch.Generator(device="cuda").manual_seed(0)
    U = torch.randn(M, M, device="cuda", generator=g).triu()
    B = torch.randn(M, T, device="cuda", generator=g)
    out = torch.empty_like(B)
    for tr in (False, True):
        r = (U.double().T if tr else U.double()) @ B.double()
        for _ in range(3):
            ops.precond_apply(U, B, out, tr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.precond_apply(U, B, out, tr)
        e1.record()
        torch.cuda.synchronize()
        err = float((out.double() - r).abs().max() / r.abs().max())
        ms = e0.elapsed_time(e1) / 20
        print("%s M=%d T=%d transposed=%d: %.3f ms  (%.0f GB/s of triangle, %.1f TFLOP/s fp64)  err %.2e"
              % (os.environ.get("ODF_PRECOND_APPLY", "own"), M, T, tr, ms, 2e-6 * M * M / ms, 1e-9 * M * M * T / ms, err), flush=True)
        del r
    del U
