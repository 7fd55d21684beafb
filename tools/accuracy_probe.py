"""Where does the score error of a C2-like fit against the fp64 oracle come from?  One sub-problem (default 200 k x 4 k,
d 1024, T 30, sigma 20, lambda 1e-3: the bench's parity sub-fit), one fp64 oracle fit on the CPU, and GPU fits that differ
in ONE arithmetic choice each: how the preconditioner is applied (explicit inverse GEMM vs triangular solves), how it is
built (tensor-core vs library), how the sweeps evaluate K (resident panels vs streamed tile), the operand kind.
    python tools/accuracy_probe.py [N M [d T sigma lambda]]        ODF_PRECOND_APPLY=cublas: round 1's fp32 sgemm applications
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "online-detection_b200"))
import odf  # noqa: E402
from oracle import falkon_oracle as orc  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def main():
    N, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200000, 4000)
    d, T, sigma, lam = 1024, 30, 20.0, 1e-3
    if len(sys.argv) > 6:
        d, T, sigma, lam = int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]), float(sys.argv[6])
    print("N=%d M=%d d=%d T=%d sigma=%g lambda=%g" % (N, M, d, T, sigma, lam), flush=True)
    torch.set_num_threads(os.cpu_count() or 1)
    X, c, Y = orc.make_synthetic(N + 8192, d, T, seed=0)
    Xt, X, Y = X[N:], X[:N].contiguous(), Y[:N].contiguous()
    C = X[orc.shared_centres(c[:N], M, seed=1)]
    t0 = time.time()
    a64 = orc.falkon_fit(X, Y, C, sigma, lam, dtype=torch.float64, eps_pc=1e-5, eps_cg=1e-7, cache_knm=True)
    s64 = orc.falkon_predict(Xt, C, a64, sigma)
    print("oracle fp64: %.0f s" % (time.time() - t0), flush=True)
    t0 = time.time()
    a32 = orc.falkon_fit(X, Y, C, sigma, lam, dtype=torch.float32, cache_knm=True)
    print("CPU fp32 port vs fp64 oracle: %.2e   (%.0f s)" % (rel(orc.falkon_predict(Xt, C, a32, sigma), s64), time.time() - t0), flush=True)
    Xg, Yg, Cg, Xtg = X.cuda(), Y.cuda(), C.cuda(), Xt.cuda()
    base = None
    for name, kw in (("default (tc build, inverse apply, resident)", {}),
                     ("library build", {"precond_build": "library"}),
                     ("trsm apply (library build)", {"precond_build": "library", "precond_apply": "trsm"}),
                     ("streamed sweeps (panel16)", {"sweep_mode": "panel16"}),
                     ("recompute sweeps", {"sweep_mode": "recompute"}),
                     ("tf32 operands", {"operand_kind": "tf32"})):
        m = odf.InCoreFalkon(kernel=odf.GaussianKernel(sigma), penalty=lam, M=M, options=odf.FalkonOptions(**kw))
        m.fit(Xg, Yg, centres=Cg)
        s = m.predict(Xtg).cpu()
        if base is None:
            base = s
        print("%-44s vs fp64 oracle %.2e   vs default %.2e   alpha vs oracle %.2e   fit %.0f ms" % (
            name, rel(s, s64), rel(s, base), rel(m.alpha_.cpu(), a64), sum(v for k, v in m.fit_times_.items() if k.endswith("_ms"))), flush=True)
    # predict alone with the ORACLE's alpha: isolates the fused tile's contribution
    m.alpha_ = a64.float().cuda()
    print("GPU predict with the oracle's alpha: %.2e" % rel(m.predict(Xtg).cpu(), s64), flush=True)


if __name__ == "__main__":
    main()
