#!/usr/bin/env python
"""CPU studies that size two round-2 items BEFORE any kernel is written.  Everything here is an
emulation in numpy / torch on the host; nothing is a GPU measurement.

    python tools/precision_study.py panel      # A: how many bits of K does a resident panel need?
    python tools/precision_study.py precond    # B: split-fp16 tensor-core GEMMs for T T^T and the inverses

A. The resident panel stores K as fp16 hi + fp16 lo (22 bits, 4 B per value).  A 3-byte panel (fp16 hi + e4m3 lo,
   ~15 bits) or a 2-byte one (fp16 hi only, 11 bits) would cut the HBM traffic of a resident sweep by 25 % / 50 %.
   The oracle's fit is run with K rounded to each format inside every sweep of the fit (predict stays exact: it runs on
   the fused tile) and the test scores are compared with the exact fit against the 1e-3 parity bar.  Also: ONE fp16
   pass for the distance product instead of three ("dist-1pass": operands rounded to fp16, exact accumulation).

B. tcgen05.mma accumulates in fp32 with TRUNCATION.  T T^T (sums of squares on the diagonal) built from 3-pass split
   fp16 operands is emulated with a truncating accumulator updated once per 16 products, for one long chain and for
   k-slices summed in round-to-nearest, and the resulting A = chol(T T^T / M + lam I) is checked for definiteness and
   for its effect on the fit."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import falkon_oracle as orc  # noqa: E402


# ---------------------------------------------------------------------------------------------- number formats
def rn_fp16(x):
    return x.to(torch.float16).to(x.dtype)


def rn_e4m3(x):
    return x.to(torch.float8_e4m3fn).to(x.dtype)


def quantise_K(K, fmt):
    """K in (0, 1] -> the value a panel of the given format holds."""
    if fmt == "exact":
        return K
    hi = rn_fp16(K)
    if fmt == "f16":                       # 2 B: hi only
        return hi
    r = (K - hi) * 4096.0                  # the panel's lo plane is scaled by 2^12
    if fmt == "f16+f16":                   # 4 B: today's panel
        return hi + rn_fp16(r) / 4096.0
    if fmt == "f16+e4m3":                  # 3 B
        return hi + rn_e4m3(r) / 4096.0
    raise ValueError(fmt)


def study_panel():
    torch.manual_seed(0)
    cases = [("C1-like  N=20000 M=1000 d=1024 T=21 sigma=15 lam=1e-3", 20000, 1000, 1024, 21, 15.0, 1e-3, 0.7),
             ("ill-cond N=20000 M=1000 d=1024 T=21 sigma=10 lam=1e-6", 20000, 1000, 1024, 21, 10.0, 1e-6, 0.7),
             ("mask-like N=30000 M=1500 d=256 T=8 sigma=10 lam=1e-6", 30000, 1500, 256, 8, 10.0, 1e-6, 0.7),
             ("tight clusters N=6000 M=500 d=1024 T=21 sigma=5 lam=1e-4 (noise 0.25)", 6000, 500, 1024, 21, 5.0, 1e-4, 0.25),
             ("RANDOM LABELS N=8000 M=2000 d=256 T=8 sigma=10 lam=1e-6 (nothing to learn: alpha is large)", 8000, 2000, 256, 8, 10.0, 1e-6, 0.7)]
    real_kernel = orc.gaussian_kernel
    for name, N, M, d, T, sigma, lam, noise in cases:
        X, c, Y = orc.make_synthetic(N, d, T, seed=0, noise=noise)
        if name.startswith("RANDOM"):
            Y = torch.sign(torch.randn(N, T, generator=torch.Generator().manual_seed(9)))
        C = X[orc.shared_centres(c, M, seed=1)]
        Xt, _, _ = orc.make_synthetic(4000, d, T, seed=11, noise=noise)
        out = {}
        for fmt in ("exact", "f16+f16", "f16+e4m3", "f16", "dist-1pass", "dist-1pass+f16"):
            if fmt.startswith("dist-1pass"):
                # ONE fp16 pass for the distance product instead of three: operands rounded to fp16 (scaled so that the
                # largest row norm sits in (128, 256]), products and sums exact; optionally K rounded to fp16 as well
                sc = 2.0 ** float(torch.floor(torch.log2(256.0 / X.norm(dim=1).max())))
                r16 = lambda t, _s=sc: (t.double() * _s).to(torch.float16).double() / _s  # noqa: E731
                kf = "f16" if fmt.endswith("+f16") else "exact"
                orc.gaussian_kernel = (lambda a, b, s, dt=torch.float64, _k=kf: quantise_K(real_kernel(r16(a), r16(b), s, dt), _k))
            else:
                orc.gaussian_kernel = (lambda a, b, s, dt=torch.float64, _f=fmt: quantise_K(real_kernel(a, b, s, dt), _f))
            # K_MM (the preconditioner) stays exact: it never goes through the panel
            pc_kernel = orc.gaussian_kernel
            try:
                t0 = time.time()
                orig_pre = orc.Preconditioner.__init__

                def pre_init(self, Cc, sg, lm, dtype=torch.float64, eps=None, kmm_dtype=torch.float64, _o=orig_pre):
                    saved = orc.gaussian_kernel
                    orc.gaussian_kernel = real_kernel
                    try:
                        _o(self, Cc, sg, lm, dtype, eps, kmm_dtype)
                    finally:
                        orc.gaussian_kernel = saved
                orc.Preconditioner.__init__ = pre_init
                alpha = orc.falkon_fit(X, Y, C, sigma, lam, dtype=torch.float64, eps_pc=1e-5, eps_cg=1e-7)
                orc.gaussian_kernel = real_kernel      # predict always runs on the fused tile (22-bit K in TMEM)
                out[fmt] = orc.falkon_predict(Xt, C, alpha, sigma)
            finally:
                orc.Preconditioner.__init__ = orig_pre
                orc.gaussian_kernel = real_kernel
            del pc_kernel
            dt = time.time() - t0
            if fmt != "exact":
                ref = out["exact"]
                rel = float((out[fmt] - ref).abs().max() / ref.abs().max())
                flips = float((out[fmt].argmax(1) != ref.argmax(1)).double().mean())
                print("  %-14s scores vs exact-K fit: max rel %.2e   argmax flips %.4f   (%.0f s)" % (fmt, rel, flips, dt), flush=True)
            else:
                print(name, flush=True)


# ---------------------------------------------------------------------------------------------- truncating GEMM
def trunc_fp32(x64):
    """float64 -> float32 rounding toward zero."""
    r = x64.astype(np.float32)
    over = np.abs(r.astype(np.float64)) > np.abs(x64)
    r[over] = np.nextafter(r[over], np.float32(0))
    return r


def split_fp16(A, scale_rows=False):
    """(hi, lo, s): s * A = hi + lo with hi, lo fp16, s a power of two (global, from the largest row norm)."""
    nrm = np.sqrt((A.astype(np.float64) ** 2).sum(1)).max()
    s = 2.0 ** np.floor(np.log2(256.0 / nrm))
    hi = (A * s).astype(np.float16)
    lo = (A * s - hi.astype(np.float64)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64), s


def gemm_nt_split_trunc(A, B, kslice=None, group=16):
    """A B^T the way the fused tile would compute it with a zero seed: 3 passes (hi.hi + hi.lo + lo.hi) of exact fp16
    products, the accumulator updated with truncation once per `group` products and pass; chains of `kslice` columns
    summed in fp32 round-to-nearest."""
    ah, al, sa = split_fp16(A)
    bh, bl, sb = split_fp16(B)
    m, k = A.shape
    n = B.shape[0]
    kslice = k if kslice is None else kslice
    total = np.zeros((m, n), dtype=np.float32)
    for k0 in range(0, k, kslice):
        acc = np.zeros((m, n), dtype=np.float32)
        for g0 in range(k0, min(k, k0 + kslice), group):
            g1 = min(k, g0 + group)
            for (x, y) in ((al, bh), (ah, bl), (ah, bh)):
                acc = trunc_fp32(acc.astype(np.float64) + x[:, g0:g1] @ y[:, g0:g1].T)
        total = (total + acc).astype(np.float32)            # fp32 round-to-nearest across slices
    return total.astype(np.float64) / (sa * sb)


def study_precond():
    torch.manual_seed(0)
    for name, M, d, sigma, lam in (("M=1536 d=256  sigma=10 lam=1e-6", 1536, 256, 10.0, 1e-6),
                                  ("M=1536 d=1024 sigma=20 lam=1e-3", 1536, 1024, 20.0, 1e-3)):
        X, c, _ = orc.make_synthetic(6 * M, d, 8, seed=0)
        C = X[orc.shared_centres(c, M, seed=1)]
        K = orc.gaussian_kernel(C, C, sigma).numpy()
        T = np.linalg.cholesky(K + 1e-5 * M * np.eye(M)).T                      # upper
        T32 = T.astype(np.float32).astype(np.float64)                           # what the GPU holds
        G_ref = T32 @ T32.T
        print(name, " cond(T) %.1e" % np.linalg.cond(T32), flush=True)
        rows = []
        for label, G in (("fp32 sgemm (RN, emulated as fp64 -> fp32)", (T32 @ T32.T).astype(np.float32).astype(np.float64)),
                         ("split fp16, one truncating chain (k = %d)" % M, gemm_nt_split_trunc(T32, T32)),
                         ("split fp16, k-slices of 512 summed in RN", gemm_nt_split_trunc(T32, T32, kslice=512)),
                         ("split fp16, k-slices of 128 summed in RN", gemm_nt_split_trunc(T32, T32, kslice=128))):
            err = G - G_ref
            dbias = float((np.diag(err) / np.diag(G_ref)).mean())
            nrm = float(np.abs(err).max() / np.abs(G_ref).max())
            Amat = G / M + lam * np.eye(M)
            mineig = float(np.linalg.eigvalsh((Amat + Amat.T) / 2).min())
            try:
                np.linalg.cholesky((Amat + Amat.T) / 2)
                pd = "PD"
            except np.linalg.LinAlgError:
                pd = "NOT positive definite"
            rows.append((label, nrm, dbias, mineig, pd))
            print("  %-46s max|err|/max|G| %.2e   mean diag bias %+.2e   min eig(A) %.3e (lam %.0e)  %s"
                  % (label, nrm, dbias, mineig, lam, pd), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("panel", "all"):
        print("== A. bits of K in a resident panel (oracle fit with K rounded inside every sweep; predict exact) ==")
        study_panel()
    if what in ("precond", "all"):
        print("== B. T T^T from split-fp16 operands with a truncating fp32 accumulator (emulation) ==")
        study_precond()
