"""Bring-up script for the GPU box: runs each check in its own process (a device-side trap kills
the CUDA context) and prints compact PASS/FAIL lines.  Not collected by pytest.

    python tests/gpu_debug.py            # all stages
    python tests/gpu_debug.py kmm mmv    # selected stages
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "online-detection_b200"))

STAGES = ["prepare", "kmm_small", "kmm", "mmv_small", "mmv", "dmmv", "panel", "precision", "precond", "cg", "perf"]   # + "bottleneck" (timing experiments, on request)


def _ref_kernel(X, C, sigma):
    import torch
    X64, C64 = X.double(), C.double()
    D = (X64 * X64).sum(1)[:, None] + (C64 * C64).sum(1)[None, :] - 2.0 * X64 @ C64.T
    return torch.exp(-D.clamp_min(0) / (2 * sigma * sigma))


def _data(n, d, seed, norm=20.0):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    X = torch.randn(n, d, generator=g)
    X = X * (norm / X.norm(dim=1).mean())
    return X.cuda()


def run_stage(stage):
    import torch
    from odf import _lib
    L = _lib.load()
    st = ctypes_stream()
    torch.manual_seed(0)
    ok = True

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

    if stage == "prepare":
        from odf import ops
        for kind in (0, 1):
            for (n, d) in [(100, 40), (1000, 1024), (257, 256)]:
                X = _data(n, d, 1)
                mean = torch.randn(d, device="cuda") * 0.1
                P = ops.Prepared(X, mean, 1.5, kind=kind)
                torch.cuda.synchronize()
                xr = ((X - mean) * 1.5).double()
                s_ = float(P.opscale[0])
                bk = 64 if kind else 32
                dp = (d + bk - 1) // bk * bk
                rec = (P.hi[:, :d].double() + P.lo[:, :d].double()) / s_
                e1 = float((rec - xr).abs().max() / xr.abs().max())
                e2 = float((P.sqn[:n].double() - (xr ** 2).sum(1)).abs().max() / 400)
                g = -0.5 * (xr ** 2).sum(1) * s_ * s_
                e3 = float(((P.hi[:, dp].double() + P.hi[:, dp + 1].double()) - g).abs().max() / g.abs().max())
                ones_ok = bool((P.hi[:, dp + bk] == 1).all() and (P.hi[:, dp + bk + 1] == 1).all() and (P.hi[:, dp + bk + 2:] == 0).all()
                               and (P.hi[:, d:dp] == 0).all() and (P.hi[:, dp + 2:dp + bk] == 0).all() and (P.sqn[n:] == 0).all())
                print(f"prepare kind={kind} n={n} d={d}: scale={s_} split_err={e1:.2e} norm_err={e2:.2e} seed_err={e3:.2e} layout_ok={ones_ok} max|hi|={float(P.hi.float().abs().max()):.1f}")
                ok &= e1 < 2e-6 and e2 < 1e-6 and e3 < 2e-6 and ones_ok
    elif stage in ("kmm_small", "kmm"):
        shapes = [(128, 32), (128, 64), (256, 32), (200, 40)] if stage == "kmm_small" else [(1000, 1024), (1500, 256), (333, 100)]
        for (M, d) in shapes:
            for sigma in (5.0, 20.0):
                C = _data(M, d, 2)
                K = torch.full((M, M), -1.0, device="cuda")
                for kind in (0, 1):
                  L.odf_set_default_kind(kind)
                  ws_b = L.odf_workspace_bytes(_lib.ODF_OP_KMM, 0, M, d, 1)
                  ws = torch.empty(ws_b, dtype=torch.uint8, device="cuda")
                  _lib.check(L.odf_gauss_kmm(_lib.ptr(C), M, d, d, sigma, _lib.ptr(K), M, _lib.ptr(ws), ws_b, st), "kmm")
                  torch.cuda.synchronize()
                  Kr = _ref_kernel(C, C, sigma)
                  err = float((K.double() - Kr).abs().max())
                  print(f"kmm kind={kind} M={M} d={d} sigma={sigma}: max_abs_err={err:.3e} diag_min={float(K.diag().min()):.7f}")
                  if err > 1e-4:
                    bad = (K.double() - Kr).abs() > 1e-4
                    idx = bad.nonzero()[:5].tolist()
                    print("   first bad idx", idx, "count", int(bad.sum()), "K", [float(K[i, j]) for i, j in idx], "ref", [float(Kr[i, j]) for i, j in idx])
                  ok &= err < (1e-4 if sigma < 10 else 1e-5)
    elif stage in ("mmv_small", "mmv"):
        shapes = [(128, 128, 32, 16), (128, 128, 32, 30), (256, 256, 64, 5), (300, 200, 40, 21)] if stage == "mmv_small" else \
                 [(5000, 1000, 1024, 21), (2000, 3000, 256, 30), (20000, 1000, 1024, 1), (777, 4500, 512, 15)]
        for (n, M, d, T) in shapes:
            sigma = 15.0
            X = _data(n, d, 3); C = _data(M, d, 4)
            V = torch.randn(M, T, device="cuda")
            out = torch.full((n, T), -7.0, device="cuda")
            ref = _ref_kernel(X, C, sigma) @ V.double()
            for kind in (0, 1):
              L.odf_set_default_kind(kind)
              ws_b = L.odf_workspace_bytes(_lib.ODF_OP_MMV, n, M, d, T)
              ws = torch.empty(ws_b, dtype=torch.uint8, device="cuda")
              _lib.check(L.odf_gauss_mmv(_lib.ptr(X), n, d, _lib.ptr(C), M, d, d, _lib.ptr(V), T, T, sigma, _lib.ptr(out), T, _lib.ptr(ws), ws_b, st), "mmv")
              torch.cuda.synchronize()
              e = rel(out, ref)
              print(f"mmv kind={kind} n={n} M={M} d={d} T={T}: rel_err={e:.3e} splits={L.odf_tile_splits(n, M, d, kind)}")
            if e > 1e-4:
                dd = (out.double() - ref).abs()
                print("   worst rows", dd.max(1).values.topk(5).indices.tolist(), "worst cols", dd.max(0).values.topk(min(5, T)).indices.tolist())
                print("   out[0,:4]", out[0, :4].tolist(), "ref[0,:4]", ref[0, :4].tolist())
            ok &= rel(out, ref) < 2e-5
    elif stage == "dmmv":
        for (n, M, d, T) in [(3000, 500, 256, 21), (20000, 1000, 1024, 30)]:
            sigma = 15.0
            X = _data(n, d, 5); C = _data(M, d, 6)
            V = torch.randn(M, T, device="cuda"); W = torch.randn(n, T, device="cuda")
            out = torch.empty(M, T, device="cuda")
            ws_b = L.odf_workspace_bytes(_lib.ODF_OP_DMMV, n, M, d, T)
            ws = torch.empty(ws_b, dtype=torch.uint8, device="cuda")
            Kr = _ref_kernel(X, C, sigma)
            for (v, w, tag) in [(V, None, "v"), (None, W, "w"), (V, W, "vw")]:
                _lib.check(L.odf_gauss_dmmv(_lib.ptr(X), n, d, _lib.ptr(C), M, d, d, _lib.ptr(v), T, _lib.ptr(w), T, T, sigma, _lib.ptr(out), T, _lib.ptr(ws), ws_b, st), "dmmv")
                torch.cuda.synchronize()
                inner = torch.zeros(n, T, dtype=torch.float64, device="cuda")
                if v is not None: inner += Kr @ v.double()
                if w is not None: inner += w.double()
                ref = Kr.T @ inner
                e = rel(out, ref)
                print(f"dmmv[{tag}] n={n} M={M} d={d} T={T}: rel_err={e:.3e}")
                ok &= e < 2e-5
    elif stage == "precond":
        for M in (300, 2000):
            d, sigma, lam, eps = 64, 5.0, 1e-4, 1e-5
            C = _data(M, d, 7)
            Kr = _ref_kernel(C, C, sigma)
            Tm = Kr.float().contiguous(); Am = torch.empty(M, M, device="cuda")
            ws_b = L.odf_workspace_bytes(_lib.ODF_OP_PRECOND, 0, M, d, 1)
            ws = torch.empty(ws_b, dtype=torch.uint8, device="cuda")
            _lib.check(L.odf_precond_init(_lib.ptr(Tm), _lib.ptr(Am), M, lam, eps, _lib.ptr(ws), ws_b, st), "precond_init")
            Tr = torch.linalg.cholesky(Kr + eps * M * torch.eye(M, device="cuda", dtype=torch.float64), upper=True)
            Ar = torch.linalg.cholesky(Tr @ Tr.T / M + lam * torch.eye(M, device="cuda", dtype=torch.float64), upper=True)
            eT, eA = rel(Tm, Tr), rel(Am, Ar)
            B = torch.randn(M, 21, device="cuda")
            errs = []
            for which, mat, tr in [(0, Tr, False), (1, Tr, True), (2, Ar, False), (3, Ar, True)]:
                Bc = B.clone()
                _lib.check(L.odf_precond_solve(_lib.ptr(Tm if which < 2 else Am), M, _lib.ptr(Bc), 21, 21, which, st), "solve")
                torch.cuda.synchronize()
                refs = torch.linalg.solve_triangular(mat.T if tr else mat, B.double(), upper=not tr)
                errs.append(rel(Bc, refs))
            from odf import ops
            Tinv = ops.precond_invert(Tm)
            Bo = torch.empty_like(B)
            ops.precond_apply(Tinv, B, Bo, False); ei1 = rel(Bo, torch.linalg.solve_triangular(Tr, B.double(), upper=True))
            ops.precond_apply(Tinv, B, Bo, True); ei2 = rel(Bo, torch.linalg.solve_triangular(Tr.T, B.double(), upper=False))
            errs += [ei1, ei2]
            print(f"precond M={M}: T_err={eT:.2e} A_err={eA:.2e} solve_errs={['%.1e' % e for e in errs]}")
            ok &= eT < 1e-3 and eA < 1e-3 and max(errs) < 5e-2
    elif stage == "cg":
        M, T = 1234, 21
        R = torch.randn(M, T, device="cuda"); P = torch.randn(M, T, device="cuda"); AP = torch.randn(M, T, device="cuda")
        Bv = torch.randn(M, T, device="cuda")
        state = torch.zeros(4 * T + 4, device="cuda")
        wsb = L.odf_cg_workspace_bytes(M, T); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        _lib.check(L.odf_cg_init(_lib.ptr(R), M, T, T, _lib.ptr(state), _lib.ptr(ws), wsb, st))
        e0 = rel(state[:T], (R.double() ** 2).sum(0))
        _lib.check(L.odf_cg_alpha(_lib.ptr(P), _lib.ptr(AP), M, T, T, 1e-7, _lib.ptr(state), _lib.ptr(ws), wsb, st))
        a_ref = (R.double() ** 2).sum(0) / ((P.double() * AP.double()).sum(0) + 1e-7)
        e1 = rel(state[T:2 * T], a_ref)
        Y = Bv.clone()
        _lib.check(L.odf_cg_axpy_a(_lib.ptr(Y), _lib.ptr(P), M, T, T, -1.0, _lib.ptr(state), st))
        e2 = rel(Y, Bv.double() - P.double() * a_ref)
        R2 = torch.randn(M, T, device="cuda")
        _lib.check(L.odf_cg_beta(_lib.ptr(R2), M, T, T, 1e-7, 1e-7, _lib.ptr(state), _lib.ptr(ws), wsb, st))
        b_ref = (R2.double() ** 2).sum(0) / ((R.double() ** 2).sum(0) + 1e-7)
        e3 = rel(state[2 * T:3 * T], b_ref)
        e3b = rel(state[:T], (R2.double() ** 2).sum(0))
        P2 = P.clone()
        _lib.check(L.odf_cg_xpby_b(_lib.ptr(P2), _lib.ptr(R2), M, T, T, _lib.ptr(state), st))
        e4 = rel(P2, R2.double() + P.double() * b_ref)
        R3 = torch.empty(M, T, device="cuda")
        _lib.check(L.odf_cg_residual(_lib.ptr(R3), _lib.ptr(Bv), _lib.ptr(AP), M, T, T, _lib.ptr(state), st))
        e5 = rel(R3, Bv.double() - AP.double())
        O = torch.empty(M, T, device="cuda")
        _lib.check(L.odf_axpby(_lib.ptr(O), 0.5, _lib.ptr(R), 2.0, _lib.ptr(P), M, T, T, st))
        e6 = rel(O, 0.5 * R.double() + 2 * P.double())
        torch.cuda.synchronize()
        print("cg errs", ["%.1e" % e for e in (e0, e1, e2, e3, e3b, e4, e5, e6)], "flag", float(state[4 * T]))
        ok &= max(e0, e1, e2, e3, e3b, e4, e5, e6) < 1e-5
    elif stage == "perf":
        from odf import ops
        for kind in (0, 1):
          for (n, M, d, T) in [(131072, 10000, 1024, 30), (262144, 5000, 256, 15), (10000, 131072, 1024, 30)]:
            sigma = 15.0
            X = _data(n, d, 8); C = _data(M, d, 9)
            V = torch.randn(M, T, device="cuda")
            px, pc = ops.Prepared(X, kind=kind), ops.Prepared(C, kind=kind)
            rhs = ops.SplitRhs(M, T, "cuda").fill(V)
            part = ops.alloc_partial(px, pc, rhs.T_pad, "cuda")
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            for it in range(4):
                if it == 1: e0.record()
                ops.mmv_partial(px, pc, rhs, sigma, part)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            fl = 2.0 * n * M * (d + T)
            print(f"perf mmv kind={kind} n={n} M={M} d={d} T={T} S={part.shape[0]}: {ms:.2f} ms  alg={fl / ms / 1e9:.1f} TFLOP/s")
    elif stage == "panel":
        from odf import ops
        for (n, M, d, T) in [(300, 200, 40, 21), (5000, 1000, 256, 30), (4096, 777, 64, 5), (131072, 10000, 1024, 30), (262144, 5000, 256, 15), (131072, 30000, 256, 21)]:
            sigma = 15.0
            X = _data(n, d, 5); C = _data(M, d, 6)
            V = torch.randn(M, T, device="cuda")
            px, pc = ops.Prepared(X), ops.Prepared(C)
            out = torch.empty(M, T, device="cuda"); out2 = torch.empty(M, T, device="cuda")
            swp = ops.Sweeper(px, pc, sigma, T, mode="panel"); swr = ops.Sweeper(px, pc, sigma, T, mode="recompute")
            swp.dmmv(V, None, out); swr.dmmv(V, None, out2)
            torch.cuda.synchronize()
            e = rel(out, out2)
            msg = f"panel n={n} M={M} d={d} T={T}: panel_vs_recompute={e:.3e}"
            ref = torch.zeros(M, T, device="cuda", dtype=torch.float64)
            for r0 in range(0, n, 8192):
                Kr = _ref_kernel(X[r0:r0 + 8192], C, sigma)
                ref += Kr.T @ (Kr @ V.double())
            msg += f" panel_vs_ref={rel(out, ref):.3e} recompute_vs_ref={rel(out2, ref):.3e}"
            ts = []
            for sw in (swp, swr):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                sw.dmmv(V, None, out); e0.record()
                for _ in range(3): sw.dmmv(V, None, out)
                e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 3)
            # time the panel kernel alone
            ops.PANEL_EVENTS = []
            swp.dmmv(V, None, out); torch.cuda.synchronize()
            pk = sum(a.elapsed_time(b) for (a, b, *_r) in ops.PANEL_EVENTS); ops.PANEL_EVENTS = None
            print(msg + f"  sweep_ms panel={ts[0]:.2f} recompute={ts[1]:.2f} panel_kernel_ms={pk:.2f} ({n * swp.ldp * 4 / pk / 1e6:.0f} GB/s panel read)")
            ok &= e < 2e-5
    elif stage == "panel16":
        from odf import ops
        shapes = [(300, 200, 40, 21), (5000, 1000, 256, 30), (4096, 777, 64, 5), (131072, 10000, 1024, 30),
                  (262144, 5000, 256, 15), (131072, 30000, 256, 21)]
        if os.environ.get("ODF_DEBUG_SMALL"):
            shapes = shapes[:3]
        for (n, M, d, T) in shapes:
            sigma = 15.0
            X = _data(n, d, 5); C = _data(M, d, 6)
            V = torch.randn(M, T, device="cuda")
            px, pc = ops.Prepared(X), ops.Prepared(C)
            ref = torch.zeros(M, T, device="cuda", dtype=torch.float64)
            for r0 in range(0, n, 8192):
                Kr = _ref_kernel(X[r0:r0 + 8192], C, sigma)
                ref += Kr.T @ (Kr @ V.double())
            msg = f"panel16 n={n} M={M} d={d} T={T}:"
            for mode in ("panel16", "panel", "recompute"):
                sw = ops.Sweeper(px, pc, sigma, T, mode=mode)
                out = torch.empty(M, T, device="cuda")
                sw.dmmv(V, None, out); torch.cuda.synchronize()
                err = rel(out, ref)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3): sw.dmmv(V, None, out)
                e1.record(); torch.cuda.synchronize()
                ops.PANEL_EVENTS = []; ops.TILE_EVENTS = []
                sw.dmmv(V, None, out); torch.cuda.synchronize()
                pk = sum(a.elapsed_time(b) for (a, b, *_r) in ops.PANEL_EVENTS)
                tk = sum(a.elapsed_time(b) for (a, b, *_r) in ops.TILE_EVENTS)
                ops.PANEL_EVENTS = None; ops.TILE_EVENTS = None
                gbs = (n * ((M + 127) // 128 * 128) * 4 / pk / 1e6) if pk > 0 else 0.0
                msg += f"\n    {mode:9s} err_vs_fp64={err:.2e} sweep_ms={e0.elapsed_time(e1) / 3:.2f} tile_ms={tk:.2f} panel_ms={pk:.2f} ({gbs:.0f} GB/s)"
                ok &= err < 2e-5
                del sw
            print(msg, flush=True)
    elif stage == "epi":
        # timing-only: which part of the tile bounds a launch (results are garbage under the debug flags)
        from odf import ops
        L2 = _lib.load()
        for (n, M, d, T) in [(131072, 30000, 256, 21), (131072, 10000, 1024, 30)]:
            X = _data(n, d, 1); C = _data(M, d, 2)
            px, pc = ops.Prepared(X, kind=1), ops.Prepared(C, kind=1)
            rhs = ops.SplitRhs(M, T, "cuda").fill(torch.randn(M, T, device="cuda"))
            part = ops.alloc_partial(px, pc, rhs.T_pad, "cuda")
            p16 = torch.empty((int(L2.odf_panel16_bytes(n, M)),), dtype=torch.uint8, device="cuda")
            for spill in (None, p16):
                res = []
                for flags in (0, 32, 10, 42, 4, 14):
                    os.environ["ODF_TILE_DEBUG"] = str(flags)
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    for it in range(3):
                        if it == 1: e0.record()
                        ops.mmv_partial(px, pc, rhs, 20.0, part, panel16=spill)
                    e1.record(); torch.cuda.synchronize()
                    res.append("%d:%.2f" % (flags, e0.elapsed_time(e1) / 2))
                os.environ["ODF_TILE_DEBUG"] = "0"
                print(f"epi n={n} M={M} d={d} T={T} spill16={spill is not None} ms by flags(1=noTMA,2=noSMMA,4=noEpiMath,8=noPV): " + " ".join(res), flush=True)
    elif stage == "bottleneck":
        # timing-only experiments (results are garbage under the debug flags)
        from odf import ops
        for kind in (1, 0):
            for (n, M, d, T) in [(131072, 10000, 1024, 30)]:
                X = _data(n, d, 1); C = _data(M, d, 2)
                px, pc = ops.Prepared(X, kind=kind), ops.Prepared(C, kind=kind)
                rhs = ops.SplitRhs(M, T, "cuda").fill(torch.randn(M, T, device="cuda"))
                part = ops.alloc_partial(px, pc, rhs.T_pad, "cuda")
                res = []
                for flags in (0,):
                    os.environ["ODF_TILE_DEBUG"] = str(flags)
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    for it in range(4):
                        if it == 1: e0.record()
                        ops.mmv_partial(px, pc, rhs, 20.0, part)
                    e1.record(); torch.cuda.synchronize()
                    res.append("%d:%.2f" % (flags, e0.elapsed_time(e1) / 3))
                os.environ["ODF_TILE_DEBUG"] = "0"
                # SM clock while the plain kernel runs back to back for ~1.5 s
                import pynvml, threading
                pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
                clk, pw, stop = [], [], threading.Event()
                def sample():
                    while not stop.is_set():
                        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
                        time.sleep(0.05)
                th = threading.Thread(target=sample); 
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                reps = 20
                for it in range(20): ops.mmv_partial(px, pc, rhs, 20.0, part)
                th.start(); e0.record()
                for it in range(reps): ops.mmv_partial(px, pc, rhs, 20.0, part)
                e1.record(); torch.cuda.synchronize(); stop.set(); th.join()
                os.environ["ODF_TILE_DEBUG"] = "16"
                ops.mmv_partial(px, pc, rhs, 20.0, part); torch.cuda.synchronize()     # in-kernel clock, hot chip
                os.environ["ODF_TILE_DEBUG"] = "0"
                time.sleep(2.0)
                os.environ["ODF_TILE_DEBUG"] = "16"
                ops.mmv_partial(px, pc, rhs, 20.0, part); torch.cuda.synchronize()     # in-kernel clock, after idling
                os.environ["ODF_TILE_DEBUG"] = "17"
                ops.mmv_partial(px, pc, rhs, 20.0, part); torch.cuda.synchronize()     # no TMA
                for fl in (17, 25, 89, 88):                              # no TMA, no PV; S-MMA N = 128, 64, 192, 256
                    os.environ["ODF_TILE_DEBUG"] = str(fl)
                    ops.mmv_partial(px, pc, rhs, 20.0, part); torch.cuda.synchronize()
                os.environ["ODF_TILE_DEBUG"] = "0"
                clk.sort(); pw.sort()
                print(f"sustained: {e0.elapsed_time(e1) / reps:.2f} ms/launch  sm_mhz median={clk[len(clk) // 2]} min={clk[0]} max={clk[-1]}  power median={pw[len(pw) // 2]:.0f} W max={pw[-1]:.0f} W")
                print(f"bottleneck kind={kind} n={n} M={M} d={d} T={T} ms by flags(1=noTMA,2=noSMMA,4=noEpi,8=noPV): " + " ".join(res))
    elif stage == "precision":
        # near-duplicate pairs: K_MM diagonal and small-distance entries, where the 3-pass product and
        # the truncating tensor-core accumulator matter most
        from odf import ops
        for kind in (0, 1):
            for (M, d, sigma) in [(1000, 1024, 5.0), (1000, 2048, 5.0), (2000, 256, 5.0)]:
                C = _data(M, d, 2)
                C[1::2] = C[0::2] + 0.01 * torch.randn(M // 2, d, device="cuda")     # near-duplicates
                K = ops.kmm(ops.Prepared(C, kind=kind), sigma)
                Kr = _ref_kernel(C, C, sigma)
                big = Kr > 1e-3
                relerr = float(((K.double() - Kr).abs() / Kr)[big].max())
                print(f"precision kind={kind} M={M} d={d} sigma={sigma}: max_rel_err_on_K>1e-3={relerr:.3e} diag_err={float((K.diag().double() - 1).abs().max()):.3e}")
                ok &= relerr < 2e-4
    print(("STAGE_PASS " if ok else "STAGE_FAIL ") + stage, flush=True)
    return ok


def ctypes_stream():
    import ctypes
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--stage":
        import torch  # noqa
        sys.exit(0 if run_stage(sys.argv[2]) else 1)
    stages = sys.argv[1:] or STAGES
    summary = []
    for s in stages:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", s], timeout=300,
                               capture_output=True, text=True)
            out = r.stdout + ("\n[stderr]\n" + r.stderr[-3000:] if r.returncode != 0 else "")
            rc = r.returncode
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            out += "\n[TIMEOUT]"
            rc = -9
        print(f"===== {s} rc={rc} ({time.time() - t0:.1f}s)\n{out}", flush=True)
        summary.append((s, rc))
    print("SUMMARY", summary)
