"""Latency anatomy of ONE refit of the reference's regime (minibootstrap: N a few thousand rows, M = 2000, d = 2048, T = 1):
wall clock of the drop-in wrapper call against the CUDA-event phases of the fit, the host share, launches per fit.
    python tools/small_fit_probe.py [N M d]
"""
import contextlib
import io
import os
import sys
import tempfile
import time

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "online-detection_b200"), os.path.join(ROOT, "online-detection_b200", "modules"),
          os.path.join(ROOT, "online-detection_b200", "modules", "region-classifier")):
    sys.path.insert(0, p)
import odf  # noqa: E402
from odf import ops  # noqa: E402
from oracle import falkon_oracle as orc  # noqa: E402


def wall(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r


def main():
    N, M, d = [int(a) for a in sys.argv[1:4]] if len(sys.argv) > 3 else (6000, 2000, 2048)
    import FALKONWrapper_with_centers_selection_incore as falkon
    cfg = {"ONLINE_REGION_CLASSIFIER": {"CLASSIFIER": {"sigma": 5.0, "lambda": 1e-4, "M": M}}}
    path = os.path.join(tempfile.mkdtemp(), "cfg.yaml")
    with open(path, "w") as fh:
        yaml.dump(cfg, fh)
    X, c, Y = orc.make_synthetic(N, d, 1, seed=0, pos_fraction=0.5)
    y = Y[:, 0].contiguous()
    Xg, yg = X.cuda(), y.cuda()
    clf = falkon.FALKONWrapper(path)
    with contextlib.redirect_stdout(io.StringIO()):
        ms_train, model = wall(lambda: clf.train(Xg, yg, sigma=5.0, lam=1e-4))
    print("FALKONWrapper.train (centre selection + fit + deepcopy): %.2f ms wall" % ms_train, flush=True)
    idx = clf.compute_indices_selection(yg)
    ms_sel, _ = wall(lambda: clf.compute_indices_selection(yg))
    C = Xg[idx]
    ms_gather, _ = wall(lambda: Xg[idx])
    print("  compute_indices_selection %.2f ms, X[indices] %.2f ms" % (ms_sel, ms_gather), flush=True)

    def fit():
        m = odf.InCoreFalkon(kernel=odf.GaussianKernel(5.0), penalty=1e-4, M=M)
        m.fit(Xg, yg, centres=C)
        return m
    l0 = ops.LAUNCHES
    ms_fit, m = wall(fit)
    per_fit = (ops.LAUNCHES - l0) / 11
    ft = m.fit_times_
    print("  InCoreFalkon.fit %.2f ms wall; CUDA-event phases: prepare %.2f precond %.2f cg %.2f (sum %.2f); %d libodf launches; sweep mode %s"
          % (ms_fit, ft["prepare_ms"], ft["precond_ms"], ft["cg_ms"], ft["prepare_ms"] + ft["precond_ms"] + ft["cg_ms"], per_fit, ft["sweep_mode"]), flush=True)
    for mode in ("recompute", "panel16", "resident"):
        def fit_m():
            mm = odf.InCoreFalkon(kernel=odf.GaussianKernel(5.0), penalty=1e-4, M=M, options=odf.FalkonOptions(sweep_mode=mode))
            mm.fit(Xg, yg, centres=C)
            return mm
        ms_m, mm = wall(fit_m)
        print("    sweep_mode=%-10s fit %.2f ms wall (cg %.2f ms)" % (mode, ms_m, mm.fit_times_["cg_ms"]), flush=True)
    import copy
    ms_copy, _ = wall(lambda: copy.deepcopy(m))
    Xs = Xg[:2000].contiguous()
    ms_pred, _ = wall(lambda: m.predict(Xs))
    print("  deepcopy(model) %.2f ms, predict(2000 rows) %.2f ms" % (ms_copy, ms_pred), flush=True)
    # host cost of the call sequence alone: time to ENQUEUE a fit (no synchronisation inside wall())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit()
    t_enq = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    print("  one fit: returned after %.2f ms (fit() synchronises at its end)" % t_enq, flush=True)


if __name__ == "__main__":
    main()
