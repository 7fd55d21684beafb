#!/usr/bin/env python
"""Bring-up / timing of odf_panel16_mmv (K . V from the fp16-plane panel, K-major A operand) on the GPU box:
spills K(X, C) with the fused tile, contracts with a wide-range V through the new kernel for both LBO/SBO readings
of the un-swizzled K-major descriptor (ODF_P16V_SWAP), and compares with the fp64 product.  Not a test."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "online-detection_b200")):
    sys.path.insert(0, p)
from odf import ops  # noqa: E402
from oracle import falkon_oracle as orc  # noqa: E402

dev = torch.device("cuda")
L = ops._lib.load()


def run(n, M, d, T, sigma=15.0, check_rows=None):
    X, _, _ = orc.make_synthetic(min(n, 20000), d, 3, seed=6)
    if n > X.shape[0]:
        g = torch.Generator().manual_seed(1)
        X = torch.cat([X] + [X[torch.randperm(X.shape[0], generator=g)] + 0.05 * torch.randn(X.shape, generator=g)
                             for _ in range(-(-n // X.shape[0]) - 1)])[:n].contiguous()
    P = X if n >= M else orc.make_synthetic(M, d, 3, seed=6)[0]
    C = P[torch.randperm(P.shape[0], generator=torch.Generator().manual_seed(7))[:M]].contiguous()
    g = torch.Generator().manual_seed(8)
    V = torch.randn(M, T, generator=g) * torch.logspace(-2, 2, T)[None, :]
    cols = ops.Prepared(C.to(dev))
    rows = ops.Prepared(X.to(dev), kind=cols.kind)
    rhs = ops.SplitRhs(M, T, dev).fill(V.to(dev))
    part1 = ops.alloc_partial(rows, cols, rhs.T_pad, dev)
    p16 = torch.empty((int(L.odf_panel16_bytes(n, M)),), dtype=torch.uint8, device=dev)
    ops.mmv_partial(rows, cols, rhs, sigma, part1, panel16=p16)
    tile_kv = part1.sum(0)[:, :T].double().cpu()                      # the tile's own K.V (same packed K values)
    Vpad = torch.zeros((1, M, rhs.T_pad), device=dev)
    Vpad[0, :, :T] = V.to(dev)
    Vf = torch.empty((M, rhs.T_pad), device=dev)
    V16 = torch.empty(((M + 127) // 128 * 128, 64), dtype=torch.float16, device=dev)
    absmax = torch.zeros(32, dtype=torch.int32, device=dev)
    ops.finish_w16(Vpad, T, Vf, absmax, V16)
    S = int(L.odf_panel16_mmv_splits(n, M))
    out_p = torch.empty((S, n, rhs.T_pad), device=dev)
    sl = slice(0, n) if check_rows is None else slice(0, check_rows)
    K = orc.gaussian_kernel(X[sl], C, sigma)
    ref = K @ V.double()
    scale = K @ V.double().abs()
    res = {}
    for swap in ((0, 1) if os.environ.get("ODF_TRY_SWAP") else (0,)):   # swap = 1 is the wrong reading (illegal address)
        os.environ["ODF_P16V_SWAP"] = str(swap)
        out_p.fill_(float("nan"))
        ops.panel16_mmv(p16, V16, absmax, n, M, out_p)
        torch.cuda.synchronize()
        out = out_p.sum(0)[:, :T].double().cpu()
        err = float(((out[sl] - ref).abs() / scale).max())
        err_tile = float(((out - tile_kv).abs() / tile_kv.abs().max(0).values).max())
        res[swap] = err
        print("  n=%d M=%d d=%d T=%d S=%d swap=%d: err vs fp64 %.3e (rel. to sum |K||V|), vs the tile's own K.V %.3e, nan=%s"
              % (n, M, d, T, S, swap, err, err_tile, bool(torch.isnan(out).any())), flush=True)
    os.environ["ODF_P16V_SWAP"] = "0"
    return rows, cols, p16, V16, absmax, out_p


def main():
    for shape in ((131, 130, 40, 16), (9000, 333, 64, 21), (20000, 1300, 96, 30), (131, 2900, 40, 16)):
        run(*shape)
    # timing at one C2 chunk: both panel kernels over the same 5.3 GB panel
    n, M, d, T = 131072, 10000, 1024, 30
    rows, cols, p16, V16, absmax, out_p = run(n, M, d, T, sigma=20.0, check_rows=2048)
    W16 = torch.zeros(((n + 127) // 128 * 128, 64), dtype=torch.float16, device=dev)
    S2 = int(L.odf_panel16_splits(n, M))
    out2 = torch.empty((S2, M, 32), device=dev)
    for name, fn in (("panel16_mmv (K v)", lambda: ops.panel16_mmv(p16, V16, absmax, n, M, out_p)),
                     ("panel16_tmm (K^T w)", lambda: ops.panel16_tmm(p16, W16, absmax, n, M, out2))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("  %s: %.3f ms per %.2f GB panel = %.0f GB/s" % (name, ms, p16.numel() / 1e9, p16.numel() / ms / 1e6), flush=True)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("done in %.1f s" % (time.time() - t0))
