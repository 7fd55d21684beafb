"""Tensor-core preconditioner build (odf_precond_build) against the library build (odf_precond_init + odf_precond_invert):
factors, inverse residuals, timing, and the effect on a fit measured against the fp64 oracle.
    python tools/precond_tc_check.py [M ...]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "online-detection_b200"))
import odf  # noqa: E402
from odf import ops  # noqa: E402
from oracle import falkon_oracle as orc  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def ev():
    return torch.cuda.Event(enable_timing=True)


def factors(M, d=256, sigma=15.0, lam=1e-5):
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(M, d, device="cuda", generator=g)
    X[M // 2:] = X[:M - M // 2] + 0.05 * torch.randn(M - M // 2, d, device="cuda", generator=g)     # near-duplicates
    X *= 20.0 / X.norm(dim=1).mean()
    K = ops.kmm(ops.Prepared(X), sigma)
    T0, A0 = ops.precond_init(K.clone(), lam, 1e-5)
    Ti0, Ai0 = ops.precond_invert(T0), ops.precond_invert(A0)
    for rep in range(3):
        Kc = K.clone()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        T1, A1, Ti1, Ai1 = ops.precond_build_tc(Kc, lam, 1e-5)
        e1.record()
        torch.cuda.synchronize()
        t_tc = e0.elapsed_time(e1)
        Kc = K.clone()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        Tl, Al = ops.precond_init(Kc, lam, 1e-5)
        ops.precond_invert(Tl), ops.precond_invert(Al)
        e1.record()
        torch.cuda.synchronize()
        print("M=%d rep %d: tensor-core build %.2f ms   library build %.2f ms" % (M, rep, t_tc, e0.elapsed_time(e1)), flush=True)
    print("   T vs library %.2e   A %.2e   Tinv %.2e   Ainv %.2e   strict lower zero: %s" %
          (rel(T1, T0), rel(A1, A0), rel(Ti1, Ti0), rel(Ai1, Ai0),
           all(float(t.tril(-1).abs().max()) == 0.0 for t in (T1, A1, Ti1, Ai1))), flush=True)
    n = min(M, 3000)
    I = torch.eye(n, device="cuda", dtype=torch.float64)
    for name, (Tx, Tix, Ax, Aix) in (("tc", (T1, Ti1, A1, Ai1)), ("library", (T0, Ti0, A0, Ai0))):
        rT = float((Tix.double()[-n:, :] @ Tx.double()[:, -n:] - I).abs().max())
        rA = float((Aix.double()[-n:, :] @ Ax.double()[:, -n:] - I).abs().max())
        Kd = K.double()[:n, :n] + 1e-5 * M * I
        bT = float(((Tx.double().T @ Tx.double())[:n, :n] - Kd).abs().max() / Kd.abs().max())
        G = (Tx.double()[:n] @ Tx.double()[:n].T) / M + lam * I
        bA = float(((Ax.double().T @ Ax.double())[:n, :n] - G).abs().max() / G.abs().max())
        print("   %-8s |Tinv T - I| %.2e  |Ainv A - I| %.2e  |T^T T - K|/|K| %.2e  |A^T A - G|/|G| %.2e" % (name, rT, rA, bT, bA), flush=True)


def fits():
    print("== fits: scores against the fp64 oracle, tensor-core vs library preconditioner ==", flush=True)
    for (N, d, T, M, sigma, lam) in ((20000, 256, 8, 2500, 15.0, 1e-4), (12000, 256, 21, 1500, 10.0, 1e-6),
                                    (20000, 1024, 21, 1000, 15.0, 1e-3)):
        X, c, Y = orc.make_synthetic(N, d, T, seed=0)
        C = X[orc.shared_centres(c, M, seed=1)]
        Xt, _, _ = orc.make_synthetic(3000, d, T, seed=11)
        t0 = time.time()
        alpha = orc.falkon_fit(X, Y, C, sigma, lam, dtype=torch.float64, eps_pc=1e-5, eps_cg=1e-7)
        ref = orc.falkon_predict(Xt, C, alpha, sigma)
        out = {}
        for build in ("tc", "library"):
            m = odf.InCoreFalkon(kernel=odf.GaussianKernel(sigma), penalty=lam, M=M, options=odf.FalkonOptions(precond_build=build))
            m.fit(X.cuda(), Y.cuda(), centres=C.cuda())
            out[build] = m.predict(Xt.cuda()).cpu().double()
            print("   N=%d d=%d T=%d M=%d sigma=%g lam=%g  %-8s rel score err vs fp64 oracle %.2e  precond %.1f ms" %
                  (N, d, T, M, sigma, lam, build, rel(out[build], ref), m.fit_times_["precond_ms"]), flush=True)
        print("      tc vs library %.2e   (oracle %.0f s)" % (rel(out["tc"], out["library"]), time.time() - t0), flush=True)


if __name__ == "__main__":
    # ODF_CHECK="d,sigma,lam" overrides the factor check's data (C3: "256,10,1e-6"); ODF_CHECK_FITS=0 skips the fits
    cfg = os.environ.get("ODF_CHECK")
    kw = {}
    if cfg:
        d_, s_, l_ = cfg.split(",")
        kw = {"d": int(d_), "sigma": float(s_), "lam": float(l_)}
    for M in [int(a) for a in sys.argv[1:]] or [2500, 10000]:
        factors(M, **kw)
    if os.environ.get("ODF_CHECK_FITS", "1") != "0":
        fits()
