"""CPU test of the host orchestration of ops.Sweeper(mode="resident") (no GPU, no compute call into libodf).

The device operators the Sweeper drives (fused tile with spill, W16 conversion, panel contraction, slab reduction)
are replaced by TEST-ONLY torch emulations that keep the same calling convention — partial slabs per column split,
"panels" identified by their buffer, W16 identified by its buffer — so the test pins the slab / chunk / orientation
bookkeeping: which panel is filled when, which one every later sweep reads, ragged last chunks, w / w_scale / scale
handling, and that all of it equals the oracle's dmmv.  The kernels themselves are checked on the GPU
(tests/test_gpu_parity.py::test_resident_*)."""
import pytest
import torch

from oracle import falkon_oracle as orc

DT = torch.float64


class _FakePrepared:
    def __init__(self, X):
        self.hi = X.to(DT).contiguous()
        self.lo = self.hi
        self.sqn = torch.zeros((X.shape[0] + 127) // 128 * 128, dtype=DT)
        self.opscale = torch.ones(2)
        self.n, self.d, self.kind, self.pitch = X.shape[0], X.shape[1], 1, X.shape[1]


_FakePreparedBase = _FakePrepared


class _FakeRhs:
    def __init__(self, m, T, device):
        self.m, self.T, self.T_pad, self.V = int(m), int(T), (16 if T <= 16 else 32), None

    def fill(self, V, scale=1.0):
        assert V.shape == (self.m, self.T)
        self.V = V.to(DT) * scale
        return self


@pytest.fixture()
def emu(monkeypatch, lib):
    from odf import ops
    store = {"panel": {}, "w16": {}, "calls": [], "hi": []}

    def tile_splits(n_rows, n_cols, d, kind):
        return min(3, (n_cols + 127) // 128)

    def alloc_partial(rows, cols, T_pad, device):
        return torch.full((tile_splits(rows.n, cols.n, rows.d, rows.kind), rows.n, T_pad), float("nan"), dtype=torch.float32)

    def mmv_partial(rows, cols, rhs, sigma, partial, panel=None, panel16=None):
        assert rhs.m == cols.n and panel is None
        K = orc.gaussian_kernel(rows.hi[:rows.n], cols.hi[:cols.n], sigma, DT)
        S = partial.shape[0]
        assert partial.shape[1] == rows.n
        edges = [cols.n * s // S for s in range(S + 1)]
        partial.zero_()
        for s in range(S):
            partial[s, :, :rhs.T] = (K[:, edges[s]:edges[s + 1]] @ rhs.V[edges[s]:edges[s + 1]]).to(torch.float32)
        if panel16 is not None:
            assert panel16.numel() >= ((rows.n + 127) // 128 * 128) * ((cols.n + 127) // 128 * 128) * 3
            store["panel"][panel16.data_ptr()] = K
        store["calls"].append(("tile", rows.n, cols.n, panel16 is not None))

    def finish_w16(partial, T, Wf, absmax, W16, addend=None):
        S, n, T_pad = partial.shape
        W = partial.to(DT).sum(0)[:, :T]
        if addend is not None:
            W = W + addend.to(DT)
        assert W16.shape[0] >= (n + 127) // 128 * 128 and Wf.shape[0] >= n
        store["w16"][W16.data_ptr()] = W
        return W16

    def panel16_tmm(panel16, W16, absmax, n_rows, M, out_partial, hi_only=False):
        store["hi"].append(("panel", hi_only))
        K = store["panel"][panel16.data_ptr()]
        W = store["w16"][W16.data_ptr()]
        assert K.shape == (n_rows, M) and W.shape[0] == n_rows, "panel read in the wrong orientation"
        S = out_partial.shape[0]
        assert out_partial.shape[1] == M
        edges = [n_rows * s // S for s in range(S + 1)]
        out_partial.zero_()
        for s in range(S):
            out_partial[s, :, :W.shape[1]] = (K[edges[s]:edges[s + 1]].T @ W[edges[s]:edges[s + 1]]).to(torch.float32)
        store["calls"].append(("panel", n_rows, M))

    def panel16_mmv(panel16, V16, absmax, n_rows, M, out_partial, hi_only=False):
        store["hi"].append(("mmv", hi_only))
        K = store["panel"][panel16.data_ptr()]
        V = store["w16"][V16.data_ptr()]
        assert K.shape == (n_rows, M) and V.shape[0] == M, "panel read in the wrong orientation"
        S = out_partial.shape[0]
        assert out_partial.shape[1] == n_rows
        edges = [M * s // S for s in range(S + 1)]
        out_partial.zero_()
        for s in range(S):
            out_partial[s, :, :V.shape[1]] = (K[:, edges[s]:edges[s + 1]] @ V[edges[s]:edges[s + 1]]).to(torch.float32)
        store["calls"].append(("mmv", n_rows, M))

    def finish_rows(partial, T, out, scale=1.0, addend=None):
        res = partial.to(DT).sum(0)[:, :T] * scale
        if addend is not None:
            res = res + addend.to(DT)
        out.copy_(res.to(out.dtype))
        return out

    for name, fn in (("tile_splits", tile_splits), ("alloc_partial", alloc_partial), ("mmv_partial", mmv_partial),
                     ("finish_w16", finish_w16), ("panel16_tmm", panel16_tmm), ("panel16_mmv", panel16_mmv),
                     ("finish_rows", finish_rows),
                     ("SplitRhs", _FakeRhs)):
        monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(ops, "PANEL_ROWS", 256)
    monkeypatch.setattr(ops, "RESIDENT_MULT", 1)
    monkeypatch.setattr(ops, "RESIDENT_SINGLE_COPY", False)
    monkeypatch.setattr(ops, "PANEL_HI_ONLY", False)
    return ops, store


@pytest.mark.parametrize("n,M,T", [(700, 150, 5), (256, 130, 21), (90, 40, 1)])
def test_resident_sweeper_bookkeeping(emu, n, M, T):
    ops, store = emu
    g = torch.Generator().manual_seed(n + M)
    X = torch.randn(n, 12, generator=g, dtype=DT)
    C = X[torch.randperm(n, generator=g)[:M]]
    sigma = 3.0
    sw = ops.Sweeper(_FakePrepared(X), _FakePrepared(C), sigma, T, mode="resident")
    assert len(sw.chunks) == -(-n // 256) and sw.chunks[-1][1] == n

    def check(v, w, scale=1.0, w_scale=1.0, tol=2e-5):
        out = torch.empty((M, T), dtype=torch.float32)
        sw.dmmv(v, w, out, scale, w_scale)
        ref = orc.dmmv(X, C, None if v is None else v.to(DT), None if w is None else w.to(DT) * w_scale, sigma, DT) * scale
        assert (out.to(DT) - ref).abs().max() <= tol * ref.abs().max()

    y = torch.randn(n, T, generator=g)
    # 1. right-hand side sweep: transposed tile pass, one per chunk, every one with the spill on; no panel pass yet
    check(None, y, w_scale=1.0 / n)
    assert store["calls"] == [("tile", M, r1 - r0, True) for (r0, r1) in sw.chunks]
    assert sw.have_tr and not sw.have_fwd
    # 2. first operator application: forward tile with spill + panel pass per chunk
    store["calls"].clear()
    check(torch.randn(M, T, generator=g), None)
    assert [c[0] for c in store["calls"]] == ["tile", "panel"] * len(sw.chunks) and sw.have_fwd
    # 3. later sweeps never evaluate a kernel value: two panel passes per chunk, K^T first read as (M x n)
    for v, w, sc in ((torch.randn(M, T, generator=g), None, 1.0), (torch.randn(M, T, generator=g), y, 0.5)):
        store["calls"].clear()
        check(v, w, scale=sc, w_scale=0.25)
        exp = []
        for (r0, r1) in sw.chunks:
            exp += [("panel", M, r1 - r0), ("panel", r1 - r0, M)]
        assert store["calls"] == exp
    # 4. K^T w alone once the forward panels are resident: one panel pass per chunk
    store["calls"].clear()
    check(None, y, scale=2.0, w_scale=0.125)
    assert store["calls"] == [("panel", r1 - r0, M) for (r0, r1) in sw.chunks]


def test_resident_sweeper_operator_first(emu):
    """An operator application before any right-hand side sweep: the transposed panels are filled on demand."""
    ops, store = emu
    g = torch.Generator().manual_seed(3)
    X = torch.randn(300, 8, generator=g, dtype=DT)
    C = X[:64]
    sw = ops.Sweeper(_FakePrepared(X), _FakePrepared(C), 2.0, 3, mode="resident")
    for _ in range(2):
        v = torch.randn(64, 3, generator=g)
        out = torch.empty((64, 3), dtype=torch.float32)
        sw.dmmv(v, None, out)
        ref = orc.dmmv(X, C, v.to(DT), None, 2.0, DT)
        assert (out.to(DT) - ref).abs().max() <= 2e-5 * ref.abs().max()
    kinds = [c[0] for c in store["calls"]]
    assert kinds == ["tile", "panel"] * 2 + ["tile"] * 2 + ["panel"] * 4


def test_resident_bytes_and_mode_names(lib):
    from odf import ops
    # C2 on one GPU: 3 B x pad(rows) x pad(centres) (x 2 orientations in the two-copy variant); 1 000 000 rows pad to
    # 1 000 064 whatever the chunking (chunks are multiples of 128 rows)
    b = ops.resident_bytes(1_000_000, 10_000)
    assert b == (1 if ops.RESIDENT_SINGLE_COPY else 2) * 3 * 10112 * 1_000_064
    with pytest.raises(ValueError):
        ops.Sweeper(_FakePrepared(torch.zeros(4, 2)), _FakePrepared(torch.zeros(2, 2)), 1.0, 1, mode="nope")


@pytest.mark.parametrize("n,M,T", [(700, 150, 5), (90, 40, 1)])
def test_resident_single_copy_bookkeeping(emu, monkeypatch, n, M, T):
    """RESIDENT_SINGLE_COPY: only K_chunk is kept; the right-hand side sweep fills it with a forward tile pass and
    K v comes from the same panel (odf_panel16_mmv)."""
    ops, store = emu
    monkeypatch.setattr(ops, "RESIDENT_SINGLE_COPY", True)
    assert ops.resident_bytes(n, M) * 2 == sum(2 * 3 * ((r + 127) // 128 * 128) * ((M + 127) // 128 * 128)
                                               for r in [min(256, n - r0) for r0 in range(0, n, 256)])
    g = torch.Generator().manual_seed(n)
    X = torch.randn(n, 12, generator=g, dtype=DT)
    C = X[torch.randperm(n, generator=g)[:M]]
    sw = ops.Sweeper(_FakePrepared(X), _FakePrepared(C), 3.0, T, mode="resident")
    assert sw.single and not hasattr(sw, "tr")

    def check(v, w, scale=1.0, w_scale=1.0):
        out = torch.empty((M, T), dtype=torch.float32)
        sw.dmmv(v, w, out, scale, w_scale)
        ref = orc.dmmv(X, C, None if v is None else v.to(DT), None if w is None else w.to(DT) * w_scale, 3.0, DT) * scale
        assert (out.to(DT) - ref).abs().max() <= 2e-5 * ref.abs().max()

    y = torch.randn(n, T, generator=g)
    check(None, y, w_scale=1.0 / n)
    assert [c[0] for c in store["calls"]] == ["tile", "panel"] * len(sw.chunks) and sw.have_fwd
    for v, w in ((torch.randn(M, T, generator=g), None), (torch.randn(M, T, generator=g), y)):
        store["calls"].clear()
        check(v, w, scale=0.5, w_scale=2.0)
        exp = []
        for (r0, r1) in sw.chunks:
            exp += [("mmv", r1 - r0, M), ("panel", r1 - r0, M)]
        assert store["calls"] == exp
    # operator first on a fresh Sweeper: forward tile pass with spill, then resident
    sw2 = ops.Sweeper(_FakePrepared(X), _FakePrepared(C), 3.0, T, mode="resident")
    store["calls"].clear()
    sw, = (sw2,)
    check(torch.randn(M, T, generator=g), y)
    check(torch.randn(M, T, generator=g), None)
    assert [c[0] for c in store["calls"]] == ["tile", "panel"] * len(sw2.chunks) + ["mmv", "panel"] * len(sw2.chunks)


@pytest.mark.parametrize("n_res", [0, 1, 2])
def test_resident_partial_bookkeeping(emu, monkeypatch, n_res):
    """Hybrid: the first n_res row chunks stay resident, the others are streamed through the transient panel and
    re-evaluate K in every sweep."""
    ops, store = emu
    monkeypatch.setattr(ops, "RESIDENT_SINGLE_COPY", True)
    n, M, T = 700, 150, 5                                          # 3 chunks: 256, 256, 188
    g = torch.Generator().manual_seed(n)
    X = torch.randn(n, 12, generator=g, dtype=DT)
    C = X[torch.randperm(n, generator=g)[:M]]
    sw = ops.Sweeper(_FakePrepared(X), _FakePrepared(C), 3.0, T, mode="resident", resident_chunks=n_res)
    assert sw.n_res == n_res and len(sw.fwd) == n_res and sw.transient is not None
    assert sw.describe() == "resident(%d of 3 row chunks, the rest streamed)" % n_res

    def check(v, w, scale=1.0, w_scale=1.0):
        out = torch.empty((M, T), dtype=torch.float32)
        sw.dmmv(v, w, out, scale, w_scale)
        ref = orc.dmmv(X, C, None if v is None else v.to(DT), None if w is None else w.to(DT) * w_scale, 3.0, DT) * scale
        assert (out.to(DT) - ref).abs().max() <= 2e-5 * ref.abs().max()

    y = torch.randn(n, T, generator=g)
    check(None, y, w_scale=1.0 / n)
    assert [c[0] for c in store["calls"]] == ["tile", "panel"] * 3
    for v, w in ((torch.randn(M, T, generator=g), None), (torch.randn(M, T, generator=g), y), (None, y)):
        store["calls"].clear()
        check(v, w, scale=0.5, w_scale=2.0)
        first = ["panel"] if v is None else ["mmv", "panel"]
        assert [c[0] for c in store["calls"]] == first * n_res + ["tile", "panel"] * (3 - n_res)


def test_resident_plan(lib, monkeypatch):
    from odf import ops
    monkeypatch.setattr(ops, "RESIDENT_SINGLE_COPY", True)
    assert ops._resident_chunk(1_000_000) == 4 * 131072 and ops._resident_chunk(1_000_000, all_resident=False) == 131072
    assert ops._resident_chunk(1000) == 1024
    per = 3 * 131072 * 10112
    assert ops.resident_plan(1_000_000, 10_000, None, budget=50e9) is None               # everything fits
    assert ops.resident_plan(1_000_000, 10_000, None, budget=4.5 * per) == 3             # transient + 3 resident
    assert ops.resident_plan(1_000_000, 10_000, None, budget=1.5 * per) == 0             # stream everything
    monkeypatch.setattr(ops, "RESIDENT_SINGLE_COPY", False)
    assert ops.resident_plan(1_000_000, 10_000, None, budget=50e9) == 0                  # two copies: 61 GB
    assert ops.resident_plan(1_000_000, 10_000, None, budget=90e9) is None


def test_auto_mode_falls_back_to_streaming_when_the_panels_do_not_fit_after_all(emu, monkeypatch):
    """mode "auto": an out-of-memory error while allocating the resident panels (plan too optimistic) is not fatal."""
    ops, store = emu
    monkeypatch.setattr(ops, "RESIDENT_SINGLE_COPY", True)
    monkeypatch.setattr(ops, "resident_plan", lambda n_rows, M, dev, budget=None: None)
    real_alloc, calls = ops.alloc_partial, {"n": 0}

    def flaky_alloc(rows, cols, T_pad, device):
        calls["n"] += 1
        if calls["n"] == 1:
            raise torch.OutOfMemoryError("simulated")
        return real_alloc(rows, cols, T_pad, device)

    monkeypatch.setattr(ops, "alloc_partial", flaky_alloc)
    g = torch.Generator().manual_seed(5)
    X = torch.randn(300, 8, generator=g, dtype=DT)
    C = X[:64]
    sw = ops.Sweeper(_FakePrepared(X), _FakePrepared(C), 2.0, 3, mode="auto")
    assert sw.mode == "panel16" and sw.describe() == "panel16" and not hasattr(sw, "fwd")
    sw.record_stream(None)                                         # walks every buffer; host tensors are skipped
    v = torch.randn(64, 3, generator=g)
    out = torch.empty((64, 3), dtype=torch.float32)
    sw.dmmv(v, None, out)
    ref = orc.dmmv(X, C, v.to(DT), None, 2.0, DT)
    assert (out.to(DT) - ref).abs().max() <= 2e-5 * ref.abs().max()
    # an explicit "resident" request does not hide the error
    calls["n"] = 0
    with pytest.raises(torch.OutOfMemoryError):
        ops.Sweeper(_FakePrepared(X), _FakePrepared(C), 2.0, 3, mode="resident")


def test_hi_only_tier_applies_to_filled_resident_panels_only(emu, monkeypatch):
    """PANEL_HI_ONLY (experimental): the passes over FILLED resident panels read the hi plane only; the pass that
    follows the tile in the same sweep (filling pass, streamed chunks) and everything in the default tier do not."""
    ops, store = emu
    monkeypatch.setattr(ops, "RESIDENT_SINGLE_COPY", True)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(700, 12, generator=g, dtype=DT)
    C = X[:150]
    y = torch.randn(700, 5, generator=g)
    out = torch.empty((150, 5), dtype=torch.float32)
    for hi in (False, True):
        monkeypatch.setattr(ops, "PANEL_HI_ONLY", hi)
        sw = ops.Sweeper(_FakePrepared(X), _FakePrepared(C), 3.0, 5, mode="resident", resident_chunks=2)   # 2 of 3 resident
        store["hi"].clear()
        sw.dmmv(None, y, out, 1.0, 1.0 / 700)                      # filling sweep: tile + panel per chunk
        assert store["hi"] == [("panel", False)] * 3
        store["hi"].clear()
        sw.dmmv(torch.randn(150, 5, generator=g), None, out)
        assert store["hi"] == [("mmv", hi), ("panel", hi)] * 2 + [("panel", False)]


@pytest.mark.parametrize("n,M,T,mode", [(700, 150, 5, "resident"), (600, 90, 3, "auto"), (700, 150, 5, "recompute")])
def test_chunked_rows_are_prepared_when_the_filling_sweep_reaches_them(emu, monkeypatch, n, M, T, mode):
    """ops.ChunkedPrepared (rows that arrive from the host chunk by chunk): the chunk-aligned, fully resident single-copy
    sweep prepares chunk i when the panel-filling pass reaches it -- in order, once, never again in later sweeps -- and
    drops the fp32 copy after the last one; any other consumer (another sweep mode, a chunk size that does not match) gets
    the whole point set.  Sweeps equal the oracle's dmmv either way."""
    ops, store = emu
    monkeypatch.setattr(ops, "RESIDENT_SINGLE_COPY", True)
    monkeypatch.setattr(ops, "resident_plan", lambda n_rows, M_, dev, budget=None: None)      # everything fits
    made = []

    class Prep(_FakePreparedBase):
        def __init__(self, X, mean=None, scale=1.0, kind=None, linear=False):
            super().__init__(X)
            made.append(int(X.shape[0]))

    monkeypatch.setattr(ops, "Prepared", Prep)
    monkeypatch.setattr(ops, "resolve_kind", lambda kind=None: 1)
    g = torch.Generator().manual_seed(n + T)
    X = torch.randn(n, 12, generator=g, dtype=DT)
    C = X[torch.randperm(n, generator=g)[:M]]
    chunk = ops._resident_chunk(n)
    n_chunks = -(-n // chunk)
    rows = ops.ChunkedPrepared(X, chunk, [None] * n_chunks)
    sw = ops.Sweeper(rows, _FakePrepared(C), 3.0, T, mode=mode)
    out = torch.empty((M, T), dtype=torch.float32)
    y = torch.randn(n, T, generator=g)
    sw.dmmv(None, y, out, 1.0, 1.0 / n)
    assert (out.to(DT) - orc.dmmv(X, C, None, y.to(DT) / n, 3.0, DT)).abs().max() <= 2e-5 * out.abs().max()
    if mode == "recompute":
        assert made == [n] and rows.parts == [None] * n_chunks            # one whole-set preparation, no chunk touched
    else:
        assert made == [min(chunk, n - r0) for r0 in range(0, n, chunk)] and rows.X is None
    for _ in range(2):
        v = torch.randn(M, T, generator=g)
        sw.dmmv(v, None, out, 0.5)
        ref = 0.5 * orc.dmmv(X, C, v.to(DT), None, 3.0, DT)
        assert (out.to(DT) - ref).abs().max() <= 2e-5 * ref.abs().max()
    assert len(made) == (1 if mode == "recompute" else n_chunks)          # later sweeps prepare nothing
    # a chunk size that is not the sweeper's: the whole set
    rows2 = ops.ChunkedPrepared(X, 128, [None] * (-(-n // 128)))
    made.clear()
    sw2 = ops.Sweeper(rows2, _FakePrepared(C), 3.0, T, mode="resident")
    sw2.dmmv(None, y, out, 1.0, 1.0 / n)
    assert made == [n] and (out.to(DT) - orc.dmmv(X, C, None, y.to(DT) / n, 3.0, DT)).abs().max() <= 2e-5 * out.abs().max()
