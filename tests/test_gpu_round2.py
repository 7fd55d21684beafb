"""GPU tests of the round-2 paths (through the C ABI; oracle / torch fp64 as the checker):
block-structured `kernel.mmv` (the *_parallel heads' alpha_parallel), z-score fused into predict, chunk-pipelined predict
of host-resident rows, the batched RLS trainer / fused apply at the sizes BASELINE config 4 names, the batched
testRegionClassifier."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import falkon_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def odf(lib):
    import odf as _odf
    return _odf


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def test_block_structured_mmv_equals_dense(odf):
    """40 classes x 130 centres stacked as the reference heads do (roi_box_predictors.py:140-160): alpha_parallel has one
    non-zero block per class.  The block-aware path (every 32-column block against its own centre rows) must reproduce
    the dense product of the oracle, for the cached second call too, with `out=` and after an in-place change of alpha."""
    from odf import ops
    d, n, Mi, Tc = 96, 3000, 130, 40
    X, _, _ = orc.make_synthetic(n, d, 3, seed=4)
    g = torch.Generator().manual_seed(5)
    C = torch.cat([X[torch.randperm(n, generator=g)[:Mi]] for _ in range(Tc)])
    alpha = torch.zeros(Tc * Mi, Tc)
    for t in range(Tc):
        alpha[t * Mi:(t + 1) * Mi, t] = torch.randn(Mi, generator=g)
    k = odf.GaussianKernel(12.0)
    Xg, Cg, ag = X.cuda(), C.cuda(), alpha.cuda()
    ranges = ops.column_block_ranges(ag)
    assert ranges == [(0, 32 * Mi), ((32 * Mi) // 128 * 128, Tc * Mi)]
    ref = orc.mmv(X, C, alpha, 12.0)
    l0 = ops.LAUNCHES
    s1 = k.mmv(Xg, Cg, ag)
    first = ops.LAUNCHES - l0
    out = torch.empty((n, Tc), device="cuda")
    l0 = ops.LAUNCHES
    s2 = k.mmv(Xg, Cg, ag, out=out)
    assert s2 is out and torch.equal(s1, s2) and ops.LAUNCHES - l0 < first          # centres and right-hand sides were cached
    assert rel(s1, ref) < 2e-5
    ag[:Mi, 0] *= 2.0                                                               # version bump: the cache must not be reused
    ref2 = orc.mmv(X, C, ag.cpu(), 12.0)
    assert rel(k.mmv(Xg, Cg, ag), ref2) < 2e-5
    # a dense right-hand side with > 32 columns keeps the full centre range for every block
    dense = torch.randn(Tc * Mi, Tc, generator=g)
    assert ops.column_block_ranges(dense.cuda()) == [(0, Tc * Mi)] * 2
    assert rel(k.mmv(Xg, Cg, dense.cuda()), orc.mmv(X, C, dense, 12.0)) < 2e-5
    assert ops.column_block_ranges(ag[:, :21]) is None


def test_predict_fused_zscore_and_pipelined_host_rows(odf, monkeypatch):
    from odf import ops
    X, c, Y = orc.make_synthetic(9000, 128, 5, seed=2)
    C = X[orc.shared_centres(c, 300, seed=1)]
    m = odf.InCoreFalkon(kernel=odf.GaussianKernel(15.0), penalty=1e-3, M=300)
    m.fit(X.cuda(), Y.cuda(), centres=C.cuda())
    raw = X * 3.0 + 0.5
    mean, scale = torch.full((128,), 0.5), 1.0 / 3.0
    a = m.predict(ops.zscore_(raw.cuda().clone(), mean.cuda(), scale))
    b = m.predict(raw.cuda(), zscore=(mean.cuda(), scale))
    assert torch.equal(a, b)
    assert rel(a, orc.falkon_predict(X, C, m.alpha_.cpu(), 15.0)) < 2e-4
    # host-resident rows: chunked upload behind the tile, result on the host.  A 2048-row chunk runs on the single-CTA tile
    # (tf32 K hi / lo in TMEM) while the 9000-row launch runs on the CTA-pair tile (fp16-packed K), and the column splits
    # depend on the rows per launch: chunked and whole scores agree to the kernels' accuracy, not bit for bit
    monkeypatch.setattr(type(m), "PREDICT_CHUNK", 2048)
    ref = orc.falkon_predict(X, C, m.alpha_.cpu(), 15.0)
    for Xh in (raw.pin_memory(), raw):
        s = m.predict(Xh, zscore=(mean.cuda(), scale))
        assert s.device.type == "cpu" and rel(s, a) < 5e-5 and rel(s, ref) < 2e-4
    assert torch.equal(m.predict(X[:1500]), m.predict(X[:1500].cuda()).cpu())            # one chunk: the same launch


def _rls_reference(X, Y, labels, classes, lam):
    """Per-class fp64 restatement on the GPU (train_region_refiner.py:25-119 via the oracle's rls_train_class)."""
    out = []
    for cid in classes:
        sel = (labels.view(-1) == cid).nonzero()[:, 0]
        out.append(None if len(sel) == 0 else orc.rls_train_class(X[sel].cpu(), Y[sel].cpu(), lam))
    return out


@pytest.mark.parametrize("d,n,n_cls", [(256, 6000, 5), (1024, 4000, 4), (2048, 3000, 3), (100, 700, 6)])
def test_batched_rls_trainer_matches_the_per_class_fp64_solve(odf, tmp_path, d, n, n_cls):
    sys.path.insert(0, os.path.join(ROOT, "online-detection_b200", "modules", "region-refiner"))
    from region_refiner_trainer import RegionRefinerTrainer
    g = torch.Generator().manual_seed(d + n)
    X = torch.randn(n, d, generator=g) * 0.7
    labels = torch.randint(1, n_cls + 1, (n, 1), generator=g).float()
    labels[labels == 2] = 3.0                                                     # class 2 has no rows
    Wtrue = torch.randn(d, 4, generator=g) * 0.05
    Y = X @ Wtrue + 0.1 * torch.randn(n, 4, generator=g) + torch.tensor([0.1, -0.2, 0.05, 0.3])
    cfg = {"CHOSEN_CLASSES": ["__background__"] + ["c%d" % i for i in range(1, n_cls + 1)]}
    lam = 10.0
    tr = RegionRefinerTrainer(cfg, lam, is_rpn=False)
    models = tr({"C": labels.cuda(), "O": None, "X": X.cuda(), "Y": Y.cuda()})
    ref = _rls_reference(X, Y, labels, range(1, n_cls + 1), lam)
    assert len(models) == n_cls
    for m, r in zip(models, ref):
        if r is None:
            assert m["Beta"] is None and m["mu"] is None
            continue
        W = torch.stack([m["Beta"][str(k)]["weights"] for k in range(4)], 1)
        Wr = torch.stack([r["Beta"][str(k)]["weights"] for k in range(4)], 1)
        assert rel(m["mu"], r["mu"]) < 1e-6 and rel(m["T"], r["T"]) < 1e-5 and rel(m["T_inv"], r["T_inv"]) < 1e-5
        assert rel(W, Wr) < 1e-5
        for k in range(4):
            assert float((m["Beta"][str(k)]["losses"].cpu() - r["Beta"][str(k)]["losses"]).abs().max()) < 1e-5


def test_fused_rls_apply_matches_torch(odf):
    from odf import ops
    g = torch.Generator().manual_seed(3)
    n, d, C = 333, 257, 7
    feat = torch.randn(n, d, generator=g).cuda()
    Wp = (torch.randn(d, 4 * C, generator=g) * 0.02).cuda()
    b = (torch.randn(4 * C, generator=g) * 0.1).cuda()
    A = torch.randn(C, 4, 4, generator=g) * 0.3 + torch.eye(4)
    Tinv, mu = A.cuda(), (torch.randn(C, 4, generator=g) * 0.1).cuda()
    x1 = torch.rand(n, generator=g) * 500
    y1 = torch.rand(n, generator=g) * 400
    ex = torch.stack((x1, y1, x1 + 10 + torch.rand(n, generator=g) * 100, y1 + 10 + torch.rand(n, generator=g) * 60), 1).cuda()
    eps = float(np.spacing(1))
    mean, zs = (torch.randn(d, generator=g) * 0.1).cuda(), 0.8
    for mean_, zs_ in ((None, 1.0), (mean, zs)):
        out = ops.rls_apply(feat, Wp, b, Tinv, mu, ex, 640, 480, eps, mean_, zs_)
        f = feat if mean_ is None else (feat - mean_) * zs_
        Yv = (f.double() @ Wp.double() + b.double()).view(n, C, 4)
        Yv = torch.einsum("nck,ckj->ncj", Yv, Tinv.double()) + mu.double()
        sw = (ex[:, 2] - ex[:, 0] + eps).double()[:, None]
        sh = (ex[:, 3] - ex[:, 1] + eps).double()[:, None]
        cx, cy = ex[:, 0:1].double() + 0.5 * sw, ex[:, 1:2].double() + 0.5 * sh
        pcx, pcy = Yv[..., 0] * sw + cx, Yv[..., 1] * sh + cy
        pw, ph = torch.exp(Yv[..., 2]) * sw, torch.exp(Yv[..., 3]) * sh
        ref = torch.stack(((pcx - 0.5 * pw).clamp(min=0), (pcy - 0.5 * ph).clamp(min=0), (pcx + 0.5 * pw - 1).clamp(max=639),
                           (pcy + 0.5 * ph - 1).clamp(max=479)), 2)
        assert out.shape == (n, C + 1, 4) and torch.equal(out[:, 0], ex)
        assert float((out[:, 1:].double() - ref).abs().max()) < 2e-3


def test_batched_test_region_classifier_matches_the_per_class_loop(odf, tmp_path):
    """testRegionClassifier scores all classes of an image with ONE kernel.mmv on the stacked centres (z-score fused):
    same scores as the reference's per-class loop of predict calls."""
    for p in ("modules", os.path.join("modules", "region-classifier")):
        sys.path.insert(0, os.path.join(ROOT, "online-detection_b200", p))
    import yaml
    import FALKONWrapper_with_centers_selection_incore as falkon
    import OnlineRegionClassifier_incore as ocr
    cfg = {"CHOSEN_CLASSES": ["__background__", "a", "b", "c"],
           "ONLINE_REGION_CLASSIFIER": {"CLASSIFIER": {"sigma": 12.0, "lambda": 1e-3, "M": 150},
                                        "MINIBOOTSTRAP": {"HARD_THRESH": -0.7, "EASY_THRESH": -0.9}}}
    path = os.path.join(str(tmp_path), "cfg.yaml")
    with open(path, "w") as fh:
        yaml.dump(cfg, fh)
    d = 64
    Xall, c, _ = orc.make_synthetic(4000, d, 3, seed=6)
    raw = Xall * 2.0 + 1.0
    stats = {"mean": torch.ones(d).cuda(), "std": torch.ones(d).cuda(), "mean_norm": torch.tensor(40.0).cuda()}
    pos = [raw[c == t + 1][:200].cuda() for t in range(3)]
    neg = [[raw[c == 0][i * 300:(i + 1) * 300].cuda() for i in range(2)] for _ in range(3)]
    torch.manual_seed(0)
    rc = ocr.OnlineRegionClassifier(falkon.FALKONWrapper(path), pos, neg, stats, cfg_path=path)
    models = rc.trainRegionClassifier()
    test = [{"gt": np.zeros(50), "boxes": np.tile(np.array([[1.0, 2.0, 30.0, 40.0]]), (50, 1)),
             "feat": raw[3000 + 50 * i:3050 + 50 * i].numpy(), "img_size": (640, 480)} for i in range(2)]
    preds = rc.testRegionClassifier(models, test)
    for i, p in enumerate(preds):
        Xz = (torch.from_numpy(test[i]["feat"]).cuda() - stats["mean"]) * (20.0 / 40.0)
        want = torch.cat([torch.full((50, 1), -1.0)] + [m.predict(Xz).cpu() for m in models], 1)
        got = p.get_field("scores")
        assert got.shape == (50, 4) and float((got - want).abs().max()) < 1e-4


@pytest.mark.parametrize("M,T", [(1, 1), (31, 3), (64, 32), (1000, 30), (2000, 1), (4097, 21), (10000, 30), (5000, 45)])
def test_triangular_apply_matches_fp64(odf, M, T):
    """odf_precond_apply / odf_precond_apply_rows (csrc/odf_tri.cu): out = U B and U^T B for an upper-triangular U whose
    strict lower part holds garbage that must not be read, against the fp64 product; an application is an exactly rounded
    linear map (fp64 accumulation), so the error is the final rounding to fp32."""
    from odf import ops
    g = torch.Generator(device="cuda").manual_seed(M * 37 + T)
    U = torch.randn(M, M, device="cuda", generator=g)
    Ut = U.triu()
    U = Ut + torch.full_like(U, float("nan")).tril(-1)          # the kernel must not touch the strict lower triangle
    B = torch.randn(M, T, device="cuda", generator=g)
    for tr in (False, True):
        ref = (Ut.double().T if tr else Ut.double()) @ B.double()
        out = torch.full_like(B, float("nan"))
        ops.precond_apply(U, B, out, tr)
        assert torch.isfinite(out).all()
        assert float((out.double() - ref).abs().max()) <= 1.01 * 2.0 ** -24 * float(ref.abs().max())
        # row ranges (the row-sharded application of a multi-GPU fit), including ragged ends
        for r0, r1 in ((0, M), (M // 3, M), (M // 5, max(M // 5 + 1, (2 * M) // 3))):
            if r1 <= r0:
                continue
            rows = torch.full((r1 - r0, T), float("nan"), device="cuda")
            ops.precond_apply_rows(U, r0, r1, B, rows, tr)
            assert torch.equal(rows, out[r0:r1]) or float((rows.double() - ref[r0:r1]).abs().max()) <= 1.01 * 2.0 ** -24 * float(ref.abs().max())


@pytest.mark.parametrize("mode", ["auto", "recompute"])
def test_chunked_upload_fit_matches_device_fit(odf, monkeypatch, mode):
    """fit() with host-resident rows spanning several resident row chunks: the rows are uploaded chunk by chunk and every
    chunk gets its split operand form (own power-of-two operand scale) when the panel-filling sweep reaches it
    (ops.ChunkedPrepared); a sweep mode that does not consume chunk by chunk gets the whole point set.  Chunk 1 holds rows
    of 1.25 x the norm of the others."""
    from odf import ops
    monkeypatch.setattr(ops, "PANEL_ROWS", 1024)
    monkeypatch.setattr(ops, "RESIDENT_MULT", 2)
    X, c, Y = orc.make_synthetic(7000, 128, 3, seed=2)
    X = X.clone()
    X[2048:4096] *= 1.25
    C = X[orc.shared_centres(c, 300, seed=1)]
    opt = odf.FalkonOptions(sweep_mode=mode)
    base = odf.InCoreFalkon(kernel=odf.GaussianKernel(15.0), penalty=1e-4, M=300, options=opt)
    base.fit(X.cuda(), Y.cuda(), centres=C.cuda())
    ref = orc.falkon_predict(X[:2000].double(), C.double(), orc.falkon_fit(X.double(), Y, C.double(), 15.0, 1e-4, dtype=torch.float64, eps_pc=1e-5, eps_cg=1e-7), 15.0)
    assert rel(base.predict(X[:2000].cuda()), ref) < 2e-4
    for Xh, Yh in ((X.pin_memory(), Y.pin_memory()), (X, Y)):
        made = []
        orig = ops.ChunkedPrepared.part
        monkeypatch.setattr(ops.ChunkedPrepared, "part", lambda self, i: (made.append(i), orig(self, i))[1])
        m = odf.InCoreFalkon(kernel=odf.GaussianKernel(15.0), penalty=1e-4, M=300, options=opt)
        m.fit(Xh, Yh, centres=C.cuda())
        monkeypatch.setattr(ops.ChunkedPrepared, "part", orig)
        assert (sorted(set(made)) == [0, 1, 2, 3]) == (mode == "auto")
        if mode == "auto":
            # same arithmetic up to the operand scales of the chunks (alpha itself is ill-conditioned: compare scores)
            assert rel(m.predict(X[:2000].cuda()), base.predict(X[:2000].cuda())) < 1e-4
        else:
            assert torch.equal(m.alpha_, base.alpha_)
        assert rel(m.predict(X[:2000].cuda()), ref) < 2e-4


@pytest.mark.parametrize("n,T,S,with_addend", [(5000, 30, 3, True), (700, 1, 1, False), (8192, 16, 2, True), (9000, 30, 3, True), (300, 21, 1, False)])
def test_finish_w16_arithmetic_restated(odf, n, T, S, with_addend):
    """odf_finish_w16: slab reduction (index order) + addend, per-column power-of-two scales from max |W|, fp16 hi | lo split --
    bit for bit the arithmetic restated here in torch (small and large, ragged, T = 1, with and without an addend)."""
    from odf import ops
    g = torch.Generator().manual_seed(n + T)
    T_pad = 16 if T <= 16 else 32
    part = torch.randn(S, n, T_pad, generator=g) * torch.logspace(-3, 3, T_pad)[None, None, :]
    add = torch.randn(n, T, generator=g) if with_addend else None
    Wf = torch.full((n, T_pad), float("nan"), device="cuda")
    W16 = torch.full(((n + 127) // 128 * 128, 64), float("nan"), dtype=torch.float16, device="cuda")
    absmax = torch.full((32,), 123, dtype=torch.int32, device="cuda")
    ops.finish_w16(part.cuda(), T, Wf, absmax, W16, None if add is None else add.cuda())
    W = torch.zeros(n, T_pad)
    for s in range(S):
        W = W + part[s]
    W = W[:, :T] + (add if add is not None else 0.0)
    assert torch.equal(Wf[:, :T].cpu(), W)
    amax = W.abs().max(0).values
    bits = amax.view(torch.int32)
    assert torch.equal(absmax[:T].cpu(), bits) and int(absmax[T:].abs().sum()) == 0
    e = ((bits >> 23) & 0xff) - 127
    scale = torch.where(amax > 0, torch.pow(2.0, (14 - e).float()), torch.ones_like(amax))
    v = W * scale[None, :]
    hi = v.half()
    lo = ((v - hi.float()) * 2048.0).half()
    got = W16.cpu()
    assert torch.equal(got[:n, :T], hi) and torch.equal(got[:n, 32:32 + T], lo)
    assert float(got[n:].abs().sum()) == 0 and float(got[:, T:32].abs().sum()) == 0 and float(got[:, 32 + T:].abs().sum()) == 0
