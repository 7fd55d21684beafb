"""Bring-up of odf_gemm_nt_split (the LINEAR store variant of the fused tile): where are the wrong entries?
    python tools/gemm_split_debug.py            # error maps for a list of shapes
    python tools/gemm_split_debug.py crash      # the shape that faulted in r2a (run under compute-sanitizer)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "online-detection_b200"))
from odf import ops  # noqa: E402


def err_map(m, n, k, sym=False, beta=0.0, kslice=None):
    g = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    A = torch.randn(m, k, generator=g)
    B = A if sym else torch.randn(n, k, generator=g)
    n = B.shape[0]
    C0 = torch.randn(m, n, generator=g)
    Ag = A.cuda()
    Bg = Ag if sym else B.cuda()
    Cg = C0.cuda() if beta != 0.0 else torch.full((m, n), float("nan"), device="cuda")
    ops.gemm_nt_split(Ag, Bg, Cg, alpha=1.0, beta=beta, kslice=kslice)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T + beta * C0.double()
    scale = A.double().abs() @ B.double().abs().T + abs(beta) * C0.double().abs()
    e = ((Cg.double().cpu() - ref).abs() / scale)
    e[torch.isnan(e)] = 9.0
    bad = e > 1e-5
    msg = "m=%d n=%d k=%d sym=%d beta=%g: max err %.2e, bad %d of %d" % (m, n, k, sym, beta, float(e.max()), int(bad.sum()), e.numel())
    if bad.any():
        # 128 x 128 tile census of the wrong entries
        tr, tc = (m + 127) // 128, (n + 127) // 128
        cen = []
        for i in range(tr):
            row = []
            for j in range(tc):
                row.append(int(bad[i * 128:(i + 1) * 128, j * 128:(j + 1) * 128].sum()))
            cen.append(row)
        msg += "\n   bad per tile (rows = row blocks): " + str(cen)
        r, c = [int(x) for x in bad.nonzero()[0]]
        msg += "\n   first bad [%d, %d]: got %.6g want %.6g" % (r, c, float(Cg[r, c]), float(ref[r, c]))
    print(msg, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "crash":
        err_map(10000, 2048, 2048)
        print("no fault", flush=True)
        sys.exit(0)
    for (m, n, k) in [(128, 128, 64), (128, 256, 64), (128, 384, 64), (128, 512, 64), (128, 640, 64), (300, 300, 64), (300, 200, 64),
                      (129, 257, 64), (129, 257, 1024), (129, 257, 2500), (384, 384, 1024), (1000, 1500, 1000), (9000, 1024, 1024)]:
        err_map(m, n, k)
    for (m, k) in [(300, 64), (384, 64), (256, 64), (1024, 1024)]:
        err_map(m, m, k, sym=True)
    err_map(300, 300, 64, beta=2.0)
    err_map(384, 384, 2048, beta=0.0)
