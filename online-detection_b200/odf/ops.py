"""Tensor-level wrappers over the libodf C ABI.  torch is used only for device memory and
streams; every arithmetic step is a libodf call.  All functions require CUDA fp32 tensors and
raise (OdfError / ValueError) otherwise — no fallback path exists."""
import ctypes
import os

import torch

from . import _lib
from ._lib import check, ptr


# Book-keeping for bench.py: number of libodf kernels launched and (optionally) CUDA-event pairs
# around every launch of the fused tile.
LAUNCHES = 0
TILE_EVENTS = None      # set to a list to record (start, stop, flops) per tile launch
PANEL_EVENTS = None     # same for the panel contraction kernel
PANEL_ROWS = 131072     # rows per spilled K panel (x pad_rows(M) x 4 B of transient workspace)
# rows per RESIDENT panel when every chunk is resident = RESIDENT_MULT x PANEL_ROWS: resident panels are not
# workspace, so a chunk may be larger -- fewer, longer launches of the panel kernels (C2 fit 0.460 -> 0.450 s at 4 x;
# profiles/r1_resident_rows_v11.log).  Partially resident fits keep PANEL_ROWS (finer residency granularity).
RESIDENT_MULT = int(os.environ.get("ODF_RESIDENT_MULT", "4"))


def _resident_chunk(n_rows, all_resident=True):
    rows = int(PANEL_ROWS) * (max(1, int(RESIDENT_MULT)) if all_resident else 1)
    return min(rows // 128 * 128, (n_rows + 127) // 128 * 128)


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    # the raw cudaStream_t of torch's current stream on the current device.  torch.cuda.current_stream() builds a Stream
    # object and re-checks the device count on every call (~10 us, a good part of the host cost of a launch in the
    # launch-bound small fits); the two C-level accessors below are what it wraps.
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def _req(t, name, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise ValueError("%s must be a torch tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must live on a CUDA device (the FALKON hot path has no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise ValueError("%s must be float32, got %s" % (name, t.dtype))
    if t.device.index != torch._C._cuda_getDevice():
        # libodf launches on the CURRENT device's current stream (one process per GPU): a tensor on another
        # device would be read through a wrong-device launch.  Wrap the call in torch.cuda.device(t.device).
        raise ValueError("%s lives on cuda:%d but the current device is cuda:%d (one device per process; use "
                         "torch.cuda.set_device / torch.cuda.device)" % (name, t.device.index, torch.cuda.current_device()))
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must be %d-D" % (name, ndim))
    return t


def _rowmajor(t):
    """Return a tensor with unit inner stride (copying only if needed) and its row pitch."""
    if t.stride(-1) != 1 or (t.dim() == 2 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t, (t.stride(0) if t.dim() == 2 else t.shape[-1])


DEFAULT_KIND = None     # None -> the library default (odf_default_kind(): fp16-scaled split)


def resolve_kind(kind=None):
    if kind is None:
        kind = DEFAULT_KIND
    if kind is None:
        return int(_lib.load().odf_default_kind())
    if isinstance(kind, str):
        return _lib.KIND_NAMES[kind.lower()]
    return int(kind)


class Prepared:
    """Split operand form of a point set (odf_prepare_points): hi/lo arrays of
    [n x odf_operand_pitch(d, kind)] tf32-in-fp32 or scaled-fp16 elements, squared norms, and the
    power-of-two operand scale on the device."""

    __slots__ = ("hi", "lo", "sqn", "opscale", "n", "d", "kind", "pitch")

    def __init__(self, X, mean=None, scale=1.0, kind=None, linear=False):
        L = _lib.load()
        X = _req(X, "X", 2)
        X, ldx = _rowmajor(X)
        self.n, self.d = int(X.shape[0]), int(X.shape[1])
        if self.n == 0:
            raise ValueError("empty point set")
        self.kind = resolve_kind(kind)
        self.pitch = int(L.odf_operand_pitch(self.d, self.kind))
        dev = X.device
        edt = torch.float16 if self.kind == _lib.ODF_KIND_F16 else torch.float32
        self.hi = torch.empty((self.n, self.pitch), dtype=edt, device=dev)
        self.lo = torch.empty((self.n, self.pitch), dtype=edt, device=dev)
        self.sqn = torch.empty((int(L.odf_pad_rows(self.n)),), dtype=torch.float32, device=dev)
        self.opscale = torch.empty((2,), dtype=torch.float32, device=dev)
        if mean is not None:
            mean = _req(mean, "mean").contiguous()
        if linear:
            # operands of the split GEMM (odf_gemm_nt_split): no z-score, zero accumulator seed
            assert mean is None and scale == 1.0
            check(L.odf_prepare_points_linear(ptr(X), self.n, self.d, ldx, self.kind, ptr(self.hi), ptr(self.lo),
                                              ptr(self.sqn), ptr(self.opscale), _stream()), "odf_prepare_points_linear")
        else:
            check(L.odf_prepare_points(ptr(X), self.n, self.d, ldx, ptr(mean), float(scale), self.kind, ptr(self.hi),
                                       ptr(self.lo), ptr(self.sqn), ptr(self.opscale), _stream()), "odf_prepare_points")
        _count(2)


class SplitRhs:
    """Transposed, split right-hand side block [T_pad x pad_rows(m)] (odf_split_rhs)."""

    __slots__ = ("hi", "lo", "ld", "T", "T_pad", "m", "hi16", "lo16", "absmax")

    def __init__(self, m, T, device):
        L = _lib.load()
        self.m, self.T = int(m), int(T)
        self.T_pad = int(L.odf_tpad(T))
        if self.T_pad < 0:
            raise ValueError("at most 32 right-hand sides per block")
        self.ld = int(L.odf_pad_rows(m))
        # tf32 hi/lo for the single-CTA tile, fp16 hi/lo (+ per-column scales) for the CTA-pair tile
        self.hi = torch.empty((self.T_pad, self.ld), dtype=torch.float32, device=device)
        self.lo = torch.empty((self.T_pad, self.ld), dtype=torch.float32, device=device)
        self.hi16 = torch.empty((self.T_pad, self.ld), dtype=torch.float16, device=device)
        self.lo16 = torch.empty((self.T_pad, self.ld), dtype=torch.float16, device=device)
        self.absmax = torch.zeros((32,), dtype=torch.int32, device=device)

    def fill(self, V, scale=1.0):
        L = _lib.load()
        V, ldv = _rowmajor(_req(V, "V", 2))
        assert V.shape[0] == self.m and V.shape[1] == self.T
        check(L.odf_split_rhs(ptr(V), self.m, self.T, ldv, float(scale), ptr(self.hi), ptr(self.lo), self.ld,
                              self.T_pad, _stream()), "odf_split_rhs")
        check(L.odf_split_rhs16(ptr(V), self.m, self.T, ldv, float(scale), ptr(self.absmax), ptr(self.hi16),
                                ptr(self.lo16), self.ld, self.T_pad, _stream()), "odf_split_rhs16")
        _count(3)
        return self


def tile_splits(n_rows, n_cols, d, kind):
    return int(_lib.load().odf_tile_splits(n_rows, n_cols, d, kind))


def alloc_partial(rows, cols, T_pad, device):
    """Partial-slab buffer for mmv_partial(rows, cols, ...)."""
    S = tile_splits(rows.n, cols.n, rows.d, rows.kind)
    return torch.empty((S, rows.n, T_pad), dtype=torch.float32, device=device)


class RowView:
    """A contiguous row range [r0, r1) of a Prepared point set (no copy)."""

    __slots__ = ("hi", "lo", "sqn", "opscale", "n", "d", "kind", "pitch")

    def __init__(self, prep, r0, r1):
        assert r0 % 128 == 0 and 0 <= r0 < r1 <= prep.n
        self.hi, self.lo, self.sqn = prep.hi[r0:r1], prep.lo[r0:r1], prep.sqn[r0:]
        self.opscale, self.n, self.d, self.kind, self.pitch = prep.opscale, r1 - r0, prep.d, prep.kind, prep.pitch


class ChunkedPrepared:
    """Rows that arrive from the HOST chunk by chunk (Falkon.fit with host-resident X): chunk i is uploaded on a side
    stream (`ready[i]` is recorded behind its copy) and turned into its split operand form -- with its OWN power-of-two
    operand scale, the tile takes the scales of rows and centres separately -- the first time a sweep asks for it, on the
    stream of that sweep.  The panel-filling sweep of a resident fit therefore starts on chunk 0 while chunk 1 is still
    on the PCIe bus.  `chunk` must be the Sweeper's resident chunk (ops._resident_chunk); any other consumer gets the
    whole point set through whole()."""

    def __init__(self, Xd, chunk, ready, mean=None, scale=1.0, kind=None):
        self.X, self.n, self.d = Xd, int(Xd.shape[0]), int(Xd.shape[1])
        self.chunk = int(chunk)
        self.bounds = [(r0, min(self.n, r0 + self.chunk)) for r0 in range(0, self.n, self.chunk)]
        assert len(ready) == len(self.bounds)
        self.ready, self.mean, self.scale = list(ready), mean, float(scale)
        self.kind = resolve_kind(kind)
        self.device = Xd.device
        self.parts = [None] * len(self.bounds)
        self._whole = None

    def part(self, i):
        if self.parts[i] is None:
            r0, r1 = self.bounds[i]
            if self.ready[i] is not None:                        # None: the rows are already there (host-logic tests)
                torch.cuda.current_stream(self.device).wait_event(self.ready[i])
            self.parts[i] = Prepared(self.X[r0:r1], self.mean, self.scale, kind=self.kind)
            if all(p is not None for p in self.parts):
                self.X = None                                    # the fp32 copy has served its purpose
        return self.parts[i]

    def views(self):
        for i in range(len(self.bounds)):
            yield self.part(i)

    def whole(self):
        """One Prepared over all rows (consumers other than the chunk-aligned resident sweep)."""
        if self._whole is None:
            if self.X is None:
                raise RuntimeError("the rows have already been consumed chunk by chunk")
            for ev in self.ready:
                if ev is not None:
                    torch.cuda.current_stream(self.device).wait_event(ev)
            self._whole = Prepared(self.X, self.mean, self.scale, kind=self.kind)
        return self._whole


class _Shape:
    """n / d / kind of a point set (what alloc_partial needs)."""
    __slots__ = ("n", "d", "kind")

    def __init__(self, n, d, kind):
        self.n, self.d, self.kind = int(n), int(d), kind


def mmv_partial(rows, cols, rhs, sigma, partial, panel=None, panel16=None):
    """partial[s] = K(rows, cols restricted to split s) @ rhs  — the fused tcgen05 tile.
    With `panel` ([rows.n x pad_rows(cols.n)] fp32) the K tiles are also spilled for panel_tmm; with `panel16`
    (uint8 buffer of odf_panel16_bytes) they are spilled as an fp16 hi plane and a one-byte residual plane (3 B per value) for panel16_tmm."""
    L = _lib.load()
    assert rows.d == cols.d and rows.kind == cols.kind and rhs.m == cols.n
    S = int(partial.shape[0])
    if panel is None and int(L.odf_tile_pair_eligible(rows.n)):
        # CTA-pair tile (cta_group::2): large launches, with or without the fp16-plane spill
        if panel16 is not None:
            assert panel16.numel() * panel16.element_size() >= int(L.odf_panel16_bytes(rows.n, cols.n))
        ev = None
        if TILE_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        check(L.odf_gauss_mmv_pair(rows.kind, ptr(rows.hi), ptr(rows.lo), ptr(rows.sqn), ptr(rows.opscale), rows.n,
                                   ptr(cols.hi), ptr(cols.lo), ptr(cols.sqn), ptr(cols.opscale), cols.n, rows.d,
                                   ptr(rhs.hi16), ptr(rhs.lo16), rhs.ld, ptr(rhs.absmax), rhs.T_pad, S, float(sigma),
                                   ptr(partial), ptr(panel16), _stream()), "odf_gauss_mmv_pair")
        if ev is not None:
            ev[1].record()
            TILE_EVENTS.append((ev[0], ev[1], rows.n, cols.n, rows.d, rhs.T))
        _count(1)
        return
    if panel16 is not None:
        assert panel is None and panel16.numel() * panel16.element_size() >= int(L.odf_panel16_bytes(rows.n, cols.n))
        ev = None
        if TILE_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        check(L.odf_gauss_mmv_prepared_spill16(rows.kind, ptr(rows.hi), ptr(rows.lo), ptr(rows.sqn), ptr(rows.opscale),
                                               rows.n, ptr(cols.hi), ptr(cols.lo), ptr(cols.sqn), ptr(cols.opscale),
                                               cols.n, rows.d, ptr(rhs.hi), ptr(rhs.lo), rhs.ld, rhs.T_pad, S,
                                               float(sigma), ptr(partial), ptr(panel16), _stream()),
              "odf_gauss_mmv_prepared_spill16")
        if ev is not None:
            ev[1].record()
            TILE_EVENTS.append((ev[0], ev[1], rows.n, cols.n, rows.d, rhs.T))
        _count(1)
        return
    if panel is not None:
        assert panel.shape[0] >= rows.n and panel.stride(1) == 1
        ev = None
        if TILE_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        check(L.odf_gauss_mmv_prepared_spill(rows.kind, ptr(rows.hi), ptr(rows.lo), ptr(rows.sqn), ptr(rows.opscale),
                                             rows.n, ptr(cols.hi), ptr(cols.lo), ptr(cols.sqn), ptr(cols.opscale),
                                             cols.n, rows.d, ptr(rhs.hi), ptr(rhs.lo), rhs.ld, rhs.T_pad, S,
                                             float(sigma), ptr(partial), ptr(panel), panel.stride(0), _stream()),
              "odf_gauss_mmv_prepared_spill")
        if ev is not None:
            ev[1].record()
            TILE_EVENTS.append((ev[0], ev[1], rows.n, cols.n, rows.d, rhs.T))
        _count(1)
        return
    ev = None
    if TILE_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    check(L.odf_gauss_mmv_prepared(rows.kind, ptr(rows.hi), ptr(rows.lo), ptr(rows.sqn), ptr(rows.opscale), rows.n,
                                   ptr(cols.hi), ptr(cols.lo), ptr(cols.sqn), ptr(cols.opscale), cols.n, rows.d,
                                   ptr(rhs.hi), ptr(rhs.lo), rhs.ld, rhs.T_pad, S, float(sigma), ptr(partial),
                                   _stream()), "odf_gauss_mmv_prepared")
    if ev is not None:
        ev[1].record()
        TILE_EVENTS.append((ev[0], ev[1], rows.n, cols.n, rows.d, rhs.T))
    _count(1)


def panel_tmm(panel, W, n_rows, M, out_partial):
    """out_partial[s] = panel[rows of split s]^T @ W   (fp32 FMA kernel streaming the panel)."""
    L = _lib.load()
    S, M_, T_pad = out_partial.shape
    assert M_ == M and W.shape[1] == T_pad and W.is_contiguous() and S == int(L.odf_panel_splits(n_rows, M))
    ev = None
    if PANEL_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    check(L.odf_panel_tmm(ptr(panel), panel.stride(0), ptr(W), n_rows, M, T_pad, S, ptr(out_partial), _stream()),
          "odf_panel_tmm")
    if ev is not None:
        ev[1].record()
        PANEL_EVENTS.append((ev[0], ev[1], n_rows, M, T_pad, "panel_tmm_kernel"))
    _count(1)


def finish_w16(partial, T, Wf, absmax, W16, addend=None):
    """W = sum_s partial[s] (+ addend) -> Wf (fp32 scratch) and its fp16 hi|lo split W16 (B operand of panel16_tmm)."""
    L = _lib.load()
    S, n, T_pad = partial.shape
    assert Wf.shape[0] >= n and Wf.shape[1] == T_pad and Wf.is_contiguous()
    assert W16.dtype == torch.float16 and W16.shape[1] == 64 and W16.shape[0] >= (n + 127) // 128 * 128 and W16.is_contiguous()
    ld_add = 0
    if addend is not None:
        addend, ld_add = _rowmajor(_req(addend, "addend", 2))
    check(L.odf_finish_w16(ptr(partial), S, n, T_pad, T, ptr(addend), ld_add, ptr(Wf), ptr(absmax), ptr(W16), _stream()),
          "odf_finish_w16")
    _count(2)
    return W16


def panel16_tmm(panel16, W16, absmax, n_rows, M, out_partial, hi_only=False):
    """out_partial[s] = K[rows of range s]^T @ W from the fp16-plane panel (tcgen05 kind::f16, HBM-streaming).
    hi_only (experimental): stream the hi plane only (K to 11 bits, half the bytes)."""
    L = _lib.load()
    S, M_, T_pad = out_partial.shape
    assert M_ == M and S == int(L.odf_panel16_splits(n_rows, M)) and out_partial.is_contiguous()
    ev = None
    if PANEL_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    fn = L.odf_panel16_tmm_hi if hi_only else L.odf_panel16_tmm
    check(fn(ptr(panel16), n_rows, M, ptr(W16), ptr(absmax), T_pad, S, ptr(out_partial), _stream()), "odf_panel16_tmm")
    if ev is not None:
        ev[1].record()
        PANEL_EVENTS.append((ev[0], ev[1], n_rows, M, T_pad, "panel16_kernel<hi>" if hi_only else "panel16_kernel"))
    _count(1)


def panel16_mmv(panel16, V16, absmax, n_rows, M, out_partial, hi_only=False):
    """out_partial[s] = K[:, column range s] @ V from the SAME fp16-plane panel (rows are the MMA's M dimension, the
    blocked planes stream through plain bulk copies) -- K v without evaluating a kernel value."""
    L = _lib.load()
    S, n_, T_pad = out_partial.shape
    assert n_ == n_rows and S == int(L.odf_panel16_mmv_splits(n_rows, M)) and out_partial.is_contiguous()
    ev = None
    if PANEL_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    fn = L.odf_panel16_mmv_hi if hi_only else L.odf_panel16_mmv
    check(fn(ptr(panel16), n_rows, M, ptr(V16), ptr(absmax), T_pad, S, ptr(out_partial), _stream()), "odf_panel16_mmv")
    if ev is not None:
        ev[1].record()
        PANEL_EVENTS.append((ev[0], ev[1], n_rows, M, T_pad, "panel16_mmv_kernel<hi>" if hi_only else "panel16_mmv_kernel"))
    _count(1)


def finish_rows(partial, T, out, scale=1.0, addend=None):
    L = _lib.load()
    S, n, T_pad = partial.shape
    assert out.stride(1) == 1
    ld_add = 0
    if addend is not None:
        addend, ld_add = _rowmajor(_req(addend, "addend", 2))
    check(L.odf_finish_rows(ptr(partial), S, n, T_pad, T, float(scale), ptr(addend), ld_add, ptr(out),
                            out.stride(0), _stream()), "odf_finish_rows")
    _count(1)
    return out


def finish_split(partial, T, rhs_out, scale=1.0, addend=None):
    L = _lib.load()
    S, n, T_pad = partial.shape
    assert rhs_out.m == n and rhs_out.T_pad == T_pad
    ld_add = 0
    if addend is not None:
        addend, ld_add = _rowmajor(_req(addend, "addend", 2))
    check(L.odf_finish_split(ptr(partial), S, n, T_pad, T, float(scale), ptr(addend), ld_add, ptr(rhs_out.hi),
                             ptr(rhs_out.lo), rhs_out.ld, _stream()), "odf_finish_split")
    _count(1)
    return rhs_out


def kmm(prep, sigma, out=None):
    L = _lib.load()
    M = prep.n
    if out is None:
        out = torch.empty((M, M), dtype=torch.float32, device=prep.hi.device)
    check(L.odf_gauss_kmm_prepared(prep.kind, ptr(prep.hi), ptr(prep.lo), ptr(prep.sqn), ptr(prep.opscale), M, prep.d,
                                   float(sigma), ptr(out), out.stride(0), _stream()), "odf_gauss_kmm_prepared")
    _count(1)
    return out


def zscore_(X, mean, scale):
    L = _lib.load()
    X = _req(X, "X", 2)
    assert X.stride(1) == 1
    mean = None if mean is None else _req(mean, "mean").contiguous()
    check(L.odf_zscore(ptr(X), X.shape[0], X.shape[1], X.stride(0), ptr(mean), float(scale), _stream()), "odf_zscore")
    _count(1)
    return X


# ---- preconditioner ---------------------------------------------------------------------------
def precond_init(Tm, lam, eps):
    """Tm: K_MM (M x M, contiguous) -> overwritten with T; returns (T, A)."""
    L = _lib.load()
    M = Tm.shape[0]
    assert Tm.is_contiguous()
    Am = torch.empty_like(Tm)
    wsb = int(L.odf_workspace_bytes(_lib.ODF_OP_PRECOND, 0, M, 0, 1))
    ws = torch.empty((wsb,), dtype=torch.uint8, device=Tm.device)
    check(L.odf_precond_init(ptr(Tm), ptr(Am), M, float(lam), float(eps), ptr(ws), wsb, _stream()), "odf_precond_init")
    _count(4)       # own kernels only (diag shifts, triangle clears); potrf/syrk are library launches
    return Tm, Am


def precond_build_tc(Kmm, lam, eps, inverses=True):
    """Tensor-core build (odf_precond_build): K_MM (M x M, contiguous, DESTROYED) -> (T, A, T^-1, A^-1), all upper
    triangular; the inverses are None when not asked for.  Raises OdfError (ODF_ERR_LINALG) on a failed pivot."""
    L = _lib.load()
    Kmm = _req(Kmm, "K_MM", 2)
    M = Kmm.shape[0]
    assert Kmm.is_contiguous() and Kmm.shape[1] == M
    Tm, Am = torch.empty_like(Kmm), torch.empty_like(Kmm)
    Ti = torch.empty_like(Kmm) if inverses else None
    Ai = torch.empty_like(Kmm) if inverses else None
    wsb = int(L.odf_precond_build_workspace_bytes(M))
    ws = torch.empty((wsb,), dtype=torch.uint8, device=Kmm.device)
    check(L.odf_precond_build(ptr(Kmm), ptr(Tm), ptr(Am), ptr(Ti), ptr(Ai), M, float(lam), float(eps), ptr(ws), wsb, _stream()),
          "odf_precond_build")
    nb = -(-M // 1024)
    _count(12 + 3 * nb * nb)     # transposes / fills / diagonal shifts + operand pre-passes and tile launches (~3 nb^2)
    return Tm, Am, Ti, Ai


def potrf_upper_(A):
    """In-place Cholesky of a symmetric row-major matrix into its upper factor (A = U^T U)."""
    L = _lib.load()
    M = A.shape[0]
    assert A.is_contiguous()
    wsb = int(L.odf_workspace_bytes(_lib.ODF_OP_PRECOND, 0, M, 0, 1))
    ws = torch.empty((wsb,), dtype=torch.uint8, device=A.device)
    check(L.odf_potrf_upper(ptr(A), M, ptr(ws), wsb, _stream()), "odf_potrf_upper")
    _count(1)
    return A


def add_diag_(A, value):
    check(_lib.load().odf_add_diag(ptr(A), A.shape[0], float(value), _stream()), "odf_add_diag")
    _count(1)
    return A


def zero_lower_(A):
    check(_lib.load().odf_zero_strict_lower(ptr(A), A.shape[0], _stream()), "odf_zero_strict_lower")
    _count(1)
    return A


def precond_solve_(Tri, B, which):
    L = _lib.load()
    assert B.stride(1) == 1
    check(L.odf_precond_solve(ptr(Tri), Tri.shape[0], ptr(B), B.shape[1], B.stride(0), int(which), _stream()),
          "odf_precond_solve")
    return B


def precond_invert(Tri):
    """Explicit inverse of an upper-triangular factor (one TRSM against the identity)."""
    L = _lib.load()
    Inv = torch.empty_like(Tri)
    check(L.odf_precond_invert(ptr(Tri), ptr(Inv), Tri.shape[0], _stream()), "odf_precond_invert")
    _count(2)
    return Inv


def precond_apply(Inv, Bin, Bout, transposed):
    """Bout = Inv @ Bin (or Inv^T @ Bin): a bandwidth-bound GEMM instead of a triangular solve."""
    L = _lib.load()
    assert Bin.stride(1) == 1 and Bout.stride(1) == 1 and Bin.stride(0) == Bout.stride(0)
    check(L.odf_precond_apply(ptr(Inv), Inv.shape[0], ptr(Bin), ptr(Bout), Bin.shape[1], Bin.stride(0),
                              1 if transposed else 0, _stream()), "odf_precond_apply")
    return Bout


GEMM_SPLIT_KSLICE = 1024    # the tensor core accumulates with truncation: contraction chains are cut here and the slices
                            # are summed in fp32 round-to-nearest by the epilogue (beta = 1); tools/precision_study.py B


def gemm_nt_split(A, B, C, alpha=1.0, beta=0.0, kslice=None, kind=None):
    """C = alpha A B^T + beta C for row-major fp32 A [m x k], B [n x k] (row-strided views allowed) on the tensor cores:
    each k-slice is split into fp16 hi/lo operands (odf_prepare_points_linear) and contracted by the fused tile with a
    linear store epilogue (odf_gemm_nt_split): 3 passes, 2 x 11 bits per operand, fp32 accumulation."""
    L = _lib.load()
    A, B, C = _req(A, "A", 2), _req(B, "B", 2), _req(C, "C", 2)
    m, k = A.shape
    n = B.shape[0]
    assert B.shape[1] == k and C.shape == (m, n) and C.stride(1) == 1
    kslice = int(GEMM_SPLIT_KSLICE if kslice is None else kslice)
    first = True
    for k0 in range(0, k, kslice):
        k1 = min(k, k0 + kslice)
        pa = Prepared(A[:, k0:k1], kind=kind, linear=True)
        pb = pa if (B is A) else Prepared(B[:, k0:k1], kind=pa.kind, linear=True)
        check(L.odf_gemm_nt_split(pa.kind, ptr(pa.hi), ptr(pa.lo), ptr(pa.sqn), ptr(pa.opscale), m, ptr(pb.hi), ptr(pb.lo),
                                  ptr(pb.sqn), ptr(pb.opscale), n, k1 - k0, float(alpha), float(beta if first else 1.0),
                                  ptr(C), C.stride(0), _stream()), "odf_gemm_nt_split")
        _count(1)
        first = False
    return C


def gemm(A, B, C, trans_a=False, trans_b=False, alpha=1.0, beta=0.0):
    """C = alpha op(A) op(B) + beta C on (possibly strided-row) fp32 views; true-fp32 cuBLAS sgemm (the library flavour of the
    row-sharded preconditioner build; the tensor-core GEMM of the default build is gemm_nt_split / odf_precond_build)."""
    L = _lib.load()
    assert A.stride(1) == 1 and B.stride(1) == 1 and C.stride(1) == 1
    m, n = C.shape
    k = A.shape[0] if trans_a else A.shape[1]
    assert (A.shape[1] if trans_a else A.shape[0]) == m and (B.shape[0] if trans_b else B.shape[1]) == n
    assert (B.shape[1] if trans_b else B.shape[0]) == k
    check(L.odf_gemm(1 if trans_a else 0, 1 if trans_b else 0, m, n, k, float(alpha), ptr(A), A.stride(0), ptr(B),
                     B.stride(0), float(beta), ptr(C), C.stride(0), _stream()), "odf_gemm")
    return C


def precond_apply_rows(Inv, r0, r1, Bin, Bout_rows, transposed):
    """Bout_rows = (op(Inv) @ Bin)[r0:r1] for an upper-triangular Inv (row block of the distributed application)."""
    L = _lib.load()
    assert Bin.stride(1) == 1 and Bout_rows.stride(1) == 1 and Bout_rows.shape[0] >= r1 - r0
    check(L.odf_precond_apply_rows(ptr(Inv), Inv.shape[0], int(r0), int(r1), ptr(Bin), ptr(Bout_rows), Bin.shape[1],
                                   Bin.stride(0), Bout_rows.stride(0), 1 if transposed else 0, _stream()),
          "odf_precond_apply_rows")
    return Bout_rows


# ---- CG vector kernels ------------------------------------------------------------------------
class CgState:
    def __init__(self, M, T, device):
        L = _lib.load()
        self.M, self.T = int(M), int(T)
        self.state = torch.zeros((4 * T + 4,), dtype=torch.float32, device=device)
        self.wsb = int(L.odf_cg_workspace_bytes(M, T))
        self.ws = torch.empty((max(self.wsb, 8),), dtype=torch.uint8, device=device)

    def init(self, R):
        check(_lib.load().odf_cg_init(ptr(R), self.M, self.T, R.stride(0), ptr(self.state), ptr(self.ws), self.wsb,
                                      _stream()), "odf_cg_init")
        _count(2)

    def alpha(self, P, AP, eps):
        check(_lib.load().odf_cg_alpha(ptr(P), ptr(AP), self.M, self.T, P.stride(0), float(eps), ptr(self.state),
                                       ptr(self.ws), self.wsb, _stream()), "odf_cg_alpha")
        _count(2)

    def axpy_a(self, Y, X, sign):
        check(_lib.load().odf_cg_axpy_a(ptr(Y), ptr(X), self.M, self.T, Y.stride(0), float(sign), ptr(self.state),
                                        _stream()), "odf_cg_axpy_a")
        _count(1)

    def residual(self, R, Bm, H):
        check(_lib.load().odf_cg_residual(ptr(R), ptr(Bm), ptr(H), self.M, self.T, R.stride(0), ptr(self.state),
                                          _stream()), "odf_cg_residual")
        _count(1)

    def beta(self, R, eps, tol):
        check(_lib.load().odf_cg_beta(ptr(R), self.M, self.T, R.stride(0), float(eps), float(tol), ptr(self.state),
                                      ptr(self.ws), self.wsb, _stream()), "odf_cg_beta")
        _count(2)

    def xpby_b(self, P, R):
        check(_lib.load().odf_cg_xpby_b(ptr(P), ptr(R), self.M, self.T, P.stride(0), ptr(self.state), _stream()),
              "odf_cg_xpby_b")
        _count(1)

    @property
    def converged_flag(self):
        return self.state[4 * self.T:4 * self.T + 1]


def axpby(out, alpha, A, beta=0.0, B=None):
    check(_lib.load().odf_axpby(ptr(out), float(alpha), ptr(A), float(beta), ptr(B), out.shape[0], out.shape[1],
                                out.stride(0), _stream()), "odf_axpby")
    _count(1)
    return out


# ---- composite operators ----------------------------------------------------------------------
def column_block_ranges(v, block=32):
    """Block structure of a right-hand side whose column blocks touch only part of the centre rows -- the `alpha_parallel`
    of the reference's *_parallel heads (roi_box_predictors.py:140-160, rpn.py:201-227, roi_mask_predictors.py:72-99): one
    non-zero row block per class.  Returns [(row_lo, row_hi)] per `block` columns (row_lo rounded down to a multiple of
    128), or None when there is a single column block (nothing to gain).  One device -> host read; callers cache it."""
    T = v.shape[1]
    if T <= block:
        return None
    nz = (v != 0)
    out = []
    idx = torch.arange(v.shape[0], device=v.device)
    big = v.shape[0]
    lo_hi = []
    for t0 in range(0, T, block):
        rows_nz = nz[:, t0:t0 + block].any(dim=1)
        lo = torch.where(rows_nz, idx, torch.full_like(idx, big)).min()
        hi = torch.where(rows_nz, idx, torch.full_like(idx, -1)).max()
        lo_hi.append(torch.stack((lo, hi)))
    for lo, hi in torch.stack(lo_hi).cpu().tolist():
        out.append((0, 0) if hi < 0 else (int(lo) // 128 * 128, int(hi) + 1))
    return out


def mmv_into(rows, cols, v, sigma, out, col_ranges=None, rhs_cache=None):
    """out = K(rows, cols) v, 32 right-hand sides per launch of the fused tile.  `col_ranges` (column_block_ranges): every
    column block is contracted against ITS centre rows only, so a block-structured v costs one evaluation of each kernel
    value in total instead of one per column block.  `rhs_cache`: a list that keeps the split right-hand sides between
    calls with the same v (filled on the first call)."""
    T = v.shape[1]
    dev = out.device
    for bi, t0 in enumerate(range(0, T, 32)):
        t1 = min(T, t0 + 32)
        lo, hi = (0, cols.n) if col_ranges is None else col_ranges[bi]
        dst = out[:, t0:t1]
        if hi <= lo:
            dst.zero_()
            continue
        cview = cols if (lo == 0 and hi == cols.n) else RowView(cols, lo, hi)
        if rhs_cache is not None and len(rhs_cache) > bi:
            rhs = rhs_cache[bi]
        else:
            rhs = SplitRhs(cview.n, t1 - t0, dev).fill(v[lo:hi, t0:t1])
            if rhs_cache is not None:
                rhs_cache.append(rhs)
        part = alloc_partial(rows, cview, rhs.T_pad, dev)
        mmv_partial(rows, cview, rhs, sigma, part)
        if out.stride(1) == 1:
            finish_rows(part, t1 - t0, dst)
        else:
            tmp = torch.empty((rows.n, t1 - t0), dtype=torch.float32, device=dev)
            finish_rows(part, t1 - t0, tmp)
            dst.copy_(tmp)


class Sweeper:
    """Pre-allocated buffers for repeated K_nm^T (K_nm V + W) sweeps with T <= 32 columns.

    mode "panel16" (default): rows go through in chunks; the fused tile computes K_chunk V and spills
    its K tiles as hi (fp16) / lo (one byte) planes to a transient panel, then the tensor-core panel kernel forms
    K_chunk^T (K_chunk V + W) streaming the panel once at HBM speed.  K is evaluated once per sweep.
    mode "panel": same with an fp32 panel and the fp32-FMA panel kernel.  mode "recompute": the second
    half re-evaluates K in the transposed orientation with the same fused tile (no panel workspace,
    2x tensor work).

    mode "resident": the fp16-plane panels of the row chunks stay in HBM for the life of the Sweeper.  Default
    (RESIDENT_SINGLE_COPY): ONE copy, K_chunk, 3 B per kernel value (30.3 GB for 1 M x 10 k, inside the 180 GB of
    a B200); the first sweep of a fit fills it as a by-product of a forward pass of the fused tile with the spill
    on, every later sweep evaluates no kernel value at all: K v comes from the panel through odf_panel16_mmv (rows
    as the MMA's M dimension) and K^T w through odf_panel16_tmm, two passes over the resident planes at HBM speed.
    mode "auto" = as many row chunks resident as fit (resident_plan), the rest streamed through one transient panel
    as in "panel16".  ODF_RESIDENT_SINGLE=0 selects the first version (K_chunk and K_chunk^T both resident, 2 x 3 B
    per value, transposed tile pass for the right-hand side sweep)."""

    def __init__(self, rows, cols, sigma, T, mode="panel16", resident_chunks=None):
        auto = mode == "auto"
        dev = cols.hi.device
        if auto:
            # as many row chunks resident as fit (single-copy variant: the rest is streamed through a transient panel)
            resident_chunks = resident_plan(rows.n, cols.n, dev)
            mode = "resident" if (resident_chunks is None or resident_chunks > 0) else "panel16"
        if isinstance(rows, ChunkedPrepared) and not (mode == "resident" and RESIDENT_SINGLE_COPY and resident_chunks is None
                                                      and rows.chunk == _resident_chunk(rows.n)):
            rows = rows.whole()                     # only the chunk-aligned, fully resident sweep consumes chunk by chunk
        try:
            self._setup(rows, cols, sigma, T, mode, resident_chunks)
        except torch.OutOfMemoryError:
            if not (auto and mode == "resident"):
                raise
            # the plan was too optimistic (fragmentation, another allocation in between): stream instead
            for k in list(self.__dict__):
                delattr(self, k)
            if dev.type == "cuda":
                torch.cuda.empty_cache()
            if isinstance(rows, ChunkedPrepared):
                rows = rows.whole()
            self._setup(rows, cols, sigma, T, "panel16", None)

    def _setup(self, rows, cols, sigma, T, mode, resident_chunks):
        L = _lib.load()
        dev = cols.hi.device
        self.rows, self.cols, self.sigma, self.T, self.mode = rows, cols, sigma, int(T), mode
        if mode not in ("panel16", "panel", "recompute", "resident"):
            raise ValueError("unknown sweep mode %r" % (mode,))
        self.v_rhs = SplitRhs(cols.n, T, dev)       # V^T  (T_pad x M)
        Tp = self.v_rhs.T_pad
        if mode != "resident":
            self.w_rhs = SplitRhs(rows.n, T, dev)   # W^T  (T_pad x n), used by "recompute" and by the v=None sweep
            self.part2 = alloc_partial(cols, rows, Tp, dev)   # rows = centres
        if mode == "panel16":
            self.chunk = min(int(PANEL_ROWS), (rows.n + 127) // 128 * 128)
            self.chunks = [(r0, min(rows.n, r0 + self.chunk)) for r0 in range(0, rows.n, self.chunk)]
            self.views = [RowView(rows, r0, r1) for (r0, r1) in self.chunks]
            self.panel16 = torch.empty((int(L.odf_panel16_bytes(self.chunk, cols.n)),), dtype=torch.uint8, device=dev)
            self.part1 = [alloc_partial(v, cols, Tp, dev) for v in self.views[:1] + self.views[-1:]]
            self.Wf = torch.empty((self.chunk, Tp), dtype=torch.float32, device=dev)
            self.W16 = torch.empty(((self.chunk + 127) // 128 * 128, 64), dtype=torch.float16, device=dev)
            self.absmax = torch.zeros((32,), dtype=torch.int32, device=dev)
            self.pslabs = [int(L.odf_panel16_splits(r1 - r0, cols.n)) for (r0, r1) in self.chunks]
            self.part3 = torch.empty((sum(self.pslabs), cols.n, Tp), dtype=torch.float32, device=dev)
        elif mode == "resident":
            M = cols.n
            self.chunk = _resident_chunk(rows.n, resident_chunks is None or not RESIDENT_SINGLE_COPY)
            self.chunks = [(r0, min(rows.n, r0 + self.chunk)) for r0 in range(0, rows.n, self.chunk)]
            if isinstance(rows, ChunkedPrepared):
                assert rows.bounds == self.chunks
                self.views = None                   # rows.views(): prepared (behind their upload) when the sweep reaches them
            else:
                self.views = [RowView(rows, r0, r1) for (r0, r1) in self.chunks]
            sizes = sorted({r1 - r0 for (r0, r1) in self.chunks})
            u8 = lambda nbytes: torch.empty((int(nbytes),), dtype=torch.uint8, device=dev)  # noqa: E731
            self.single = bool(RESIDENT_SINGLE_COPY)
            # the first n_res chunks stay resident; the others (single-copy variant only) go through a transient
            # panel and re-evaluate K in every sweep, as in mode "panel16"
            self.n_res = len(self.chunks) if (resident_chunks is None or not self.single) else \
                max(0, min(int(resident_chunks), len(self.chunks)))
            self.fwd = [u8(L.odf_panel16_bytes(r1 - r0, M)) for (r0, r1) in self.chunks[:self.n_res]]     # K_chunk
            self.transient = u8(L.odf_panel16_bytes(self.chunk, M)) if self.n_res < len(self.chunks) else None
            self.have_fwd = self.have_tr = False
            self.part1 = {n: alloc_partial(_Shape(n, rows.d, rows.kind), cols, Tp, dev) for n in sizes}
            if self.single:
                kv_splits = {n: int(L.odf_panel16_mmv_splits(n, M)) for n in sizes}
            else:
                self.tr = [u8(L.odf_panel16_bytes(M, r1 - r0)) for (r0, r1) in self.chunks]  # K_chunk^T
                self.w_rhs_c = {n: SplitRhs(n, T, dev) for n in sizes}
                kv_splits = {n: int(L.odf_panel16_splits(M, n)) for n in sizes}
            self.kv_part = {n: torch.empty((kv_splits[n], n, Tp), dtype=torch.float32, device=dev) for n in sizes}
            self.Wf = torch.empty((self.chunk, Tp), dtype=torch.float32, device=dev)
            self.W16 = torch.empty(((self.chunk + 127) // 128 * 128, 64), dtype=torch.float16, device=dev)
            self.Wpad = torch.zeros((1, self.chunk, Tp), dtype=torch.float32, device=dev)   # padded columns stay 0
            self.absmax = torch.zeros((32,), dtype=torch.int32, device=dev)
            self.Vpad = torch.zeros((1, M, Tp), dtype=torch.float32, device=dev)
            self.Vf = torch.empty((M, Tp), dtype=torch.float32, device=dev)
            self.V16 = torch.empty(((M + 127) // 128 * 128, 64), dtype=torch.float16, device=dev)
            self.absmax_v = torch.zeros((32,), dtype=torch.int32, device=dev)
            self.pslabs = [int(L.odf_panel16_splits(r1 - r0, M)) for (r0, r1) in self.chunks]
            self.part3 = torch.empty((sum(self.pslabs), M, Tp), dtype=torch.float32, device=dev)
        elif mode == "panel":
            self.chunk = min(int(PANEL_ROWS), (rows.n + 127) // 128 * 128)
            self.chunks = [(r0, min(rows.n, r0 + self.chunk)) for r0 in range(0, rows.n, self.chunk)]
            self.views = [RowView(rows, r0, r1) for (r0, r1) in self.chunks]
            self.ldp = int(L.odf_pad_rows(cols.n))
            self.panel = torch.empty((self.chunk, self.ldp), dtype=torch.float32, device=dev)
            self.part1 = [alloc_partial(v, cols, Tp, dev) for v in self.views[:1] + self.views[-1:]]
            self.Wc = torch.zeros((self.chunk, Tp), dtype=torch.float32, device=dev)   # padded columns stay 0
            self.pslabs = [int(L.odf_panel_splits(r1 - r0, cols.n)) for (r0, r1) in self.chunks]
            self.part3 = torch.empty((sum(self.pslabs), cols.n, Tp), dtype=torch.float32, device=dev)
        else:
            self.part1 = alloc_partial(rows, cols, Tp, dev)   # rows = data
            self.Wfull = torch.empty((rows.n, int(T)), dtype=torch.float32, device=dev)

    def dmmv(self, v, w, out, scale=1.0, w_scale=1.0):
        """out = scale * K^T (K v + w_scale * w)   (local rows only; caller all-reduces)."""
        if self.mode == "resident":
            return self._dmmv_resident(v, w, out, scale, w_scale)
        if v is None:
            self.w_rhs.fill(w, w_scale)
            mmv_partial(self.cols, self.rows, self.w_rhs, self.sigma, self.part2)
            return finish_rows(self.part2, self.T, out, scale)
        if w is not None and w_scale != 1.0:
            w = w * w_scale
        self.v_rhs.fill(v)
        if self.mode == "panel16":
            slab = 0
            for i, ((r0, r1), view) in enumerate(zip(self.chunks, self.views)):
                n = r1 - r0
                part1 = self.part1[0] if n == self.part1[0].shape[1] else self.part1[-1]
                mmv_partial(view, self.cols, self.v_rhs, self.sigma, part1, panel16=self.panel16)
                finish_w16(part1, self.T, self.Wf, self.absmax, self.W16, None if w is None else w[r0:r1])
                S = self.pslabs[i]
                panel16_tmm(self.panel16, self.W16, self.absmax, n, self.cols.n, self.part3[slab:slab + S])
                slab += S
            return finish_rows(self.part3, self.T, out, scale)
        if self.mode != "panel":
            mmv_partial(self.rows, self.cols, self.v_rhs, self.sigma, self.part1)
            finish_rows(self.part1, self.T, self.Wfull, 1.0, w)
            self.w_rhs.fill(self.Wfull)          # both operand formats (tf32 for the single-CTA tile, fp16 for the pair)
            mmv_partial(self.cols, self.rows, self.w_rhs, self.sigma, self.part2)
            return finish_rows(self.part2, self.T, out, scale)
        slab = 0
        for i, ((r0, r1), view) in enumerate(zip(self.chunks, self.views)):
            n = r1 - r0
            part1 = self.part1[0] if n == self.part1[0].shape[1] else self.part1[-1]
            mmv_partial(view, self.cols, self.v_rhs, self.sigma, part1, panel=self.panel)
            finish_rows(part1, self.T, self.Wc[:n], 1.0, None if w is None else w[r0:r1])
            S = self.pslabs[i]
            panel_tmm(self.panel, self.Wc[:n], n, self.cols.n, self.part3[slab:slab + S])
            slab += S
        return finish_rows(self.part3, self.T, out, scale)

    # ---- mode "resident" ------------------------------------------------------------------------
    def _fill_transposed(self, w=None, w_scale=1.0, out=None, scale=1.0):
        """One pass of the fused tile in the transposed orientation (rows = centres, columns = one row chunk
        of X) with the spill on: leaves K_chunk^T resident and, as a by-product, out = scale * K^T (w_scale w)."""
        M, dev = self.cols.n, self.cols.hi.device
        Tp = self.v_rhs.T_pad
        S_c = [tile_splits(M, r1 - r0, self.cols.d, self.cols.kind) for (r0, r1) in self.chunks]
        partT = torch.empty((sum(S_c), M, Tp), dtype=torch.float32, device=dev)
        slab = 0
        for i, ((r0, r1), view) in enumerate(zip(self.chunks, self.views)):
            rhs = self.w_rhs_c[r1 - r0]
            if w is None:
                rhs.fill(self.Wpad[0, :r1 - r0, :self.T])           # zeros: only the spill matters
            else:
                rhs.fill(w[r0:r1], w_scale)
            mmv_partial(self.cols, view, rhs, self.sigma, partT[slab:slab + S_c[i]], panel16=self.tr[i])
            slab += S_c[i]
        self.have_tr = True
        if out is not None:
            finish_rows(partT, self.T, out, scale)
        return out

    def _dmmv_single(self, v, w, out, scale, w_scale):
        """Single-copy resident sweep.  Per row chunk: a RESIDENT chunk evaluates K once (the first sweep runs the
        fused tile with the spill into the chunk's own panel; K v is its by-product) and afterwards costs two panel
        passes, odf_panel16_mmv (K v) and odf_panel16_tmm (K^T (K v + w)); a STREAMED chunk (beyond n_res) runs the
        tile with the spill into the transient panel in every sweep."""
        M, T = self.cols.n, self.T
        fill = not self.have_fwd
        hi = bool(PANEL_HI_ONLY)            # experimental: passes over filled resident panels read the hi plane only
        need_tile = fill or self.n_res < len(self.chunks)
        if v is None:
            if need_tile:
                self.Vpad.zero_()
                self.v_rhs.fill(self.Vpad[0, :, :T])                # the tile's K.0 by-product is dropped
        else:
            if w is not None and w_scale != 1.0:
                w = w * w_scale
            if need_tile:
                self.v_rhs.fill(v)
            if not fill and self.n_res > 0:
                self.Vpad[0, :, :T].copy_(v)
                finish_w16(self.Vpad, T, self.Vf, self.absmax_v, self.V16)
        slab = 0
        part3 = self.part3
        views = self.views if self.views is not None else (self.rows.views() if need_tile else [None] * len(self.chunks))
        for i, ((r0, r1), view) in enumerate(zip(self.chunks, views)):
            n = r1 - r0
            resident = i < self.n_res
            panel = self.fwd[i] if resident else self.transient
            tile = fill or not resident
            if tile:
                mmv_partial(view, self.cols, self.v_rhs, self.sigma, self.part1[n], panel16=panel)
            if v is None:
                self.Wpad[0, :n, :T].copy_(w[r0:r1])
                if w_scale != 1.0:
                    self.Wpad[0, :n, :T].mul_(w_scale)
                finish_w16(self.Wpad[:, :n], T, self.Wf, self.absmax, self.W16)
            elif tile:
                finish_w16(self.part1[n], T, self.Wf, self.absmax, self.W16, None if w is None else w[r0:r1])
            else:
                kv = self.kv_part[n]
                panel16_mmv(panel, self.V16, self.absmax_v, n, M, kv, hi_only=hi)           # K_chunk v, same panel
                finish_w16(kv, T, self.Wf, self.absmax, self.W16, None if w is None else w[r0:r1])
            S = self.pslabs[i]
            panel16_tmm(panel, self.W16, self.absmax, n, M, part3[slab:slab + S], hi_only=hi and not tile)  # K_chunk^T (K_chunk v + w)
            slab += S
        self.have_fwd = True
        if self.n_res == len(self.chunks):
            self.part1 = None                                       # only the filling pass needs the tile's slabs
        return finish_rows(part3[:slab], T, out, scale)

    def record_stream(self, stream):
        """Mark every buffer of the Sweeper as in use on `stream` (torch.Tensor.record_stream): needed when a sweep is
        launched on a stream other than the one the Sweeper was built on, because buffers the sweep drops (the tile's
        slab buffer after the panel-filling pass) would otherwise be handed out again while that sweep still runs."""
        def walk(v):
            if isinstance(v, torch.Tensor):
                if v.is_cuda:
                    v.record_stream(stream)
            elif isinstance(v, dict):
                for x in v.values():
                    walk(x)
            elif isinstance(v, (list, tuple)):
                for x in v:
                    walk(x)
            elif isinstance(v, SplitRhs):
                for name in (getattr(type(v), "__slots__", None) or list(getattr(v, "__dict__", {}))):
                    walk(getattr(v, name, None))
        for v in self.__dict__.values():
            walk(v)

    def describe(self):
        if self.mode != "resident":
            return self.mode
        if self.n_res == len(self.chunks):
            return "resident"
        return "resident(%d of %d row chunks, the rest streamed)" % (self.n_res, len(self.chunks))

    def _dmmv_resident(self, v, w, out, scale, w_scale):
        if self.single:
            return self._dmmv_single(v, w, out, scale, w_scale)
        # two-copy variant (ODF_RESIDENT_SINGLE=0): K_chunk and K_chunk^T both resident, only odf_panel16_tmm is used
        M = self.cols.n
        if v is None:
            if not self.have_fwd:
                # right-hand side sweep of a fit: K^T (w_scale w) by the transposed tile, K_chunk^T stays resident
                return self._fill_transposed(w, w_scale, out, scale)
            slab = 0
            for i, (r0, r1) in enumerate(self.chunks):
                n = r1 - r0
                self.Wpad[0, :n, :self.T].copy_(w[r0:r1])
                if w_scale != 1.0:
                    self.Wpad[0, :n, :self.T].mul_(w_scale)
                finish_w16(self.Wpad[:, :n], self.T, self.Wf, self.absmax, self.W16)
                S = self.pslabs[i]
                panel16_tmm(self.fwd[i], self.W16, self.absmax, n, M, self.part3[slab:slab + S])
                slab += S
            return finish_rows(self.part3, self.T, out, scale)
        if w is not None and w_scale != 1.0:
            w = w * w_scale
        if not self.have_fwd:
            # first operator application: forward tile with the spill on (K v is its by-product), K_chunk stays resident
            self.v_rhs.fill(v)
            slab = 0
            for i, ((r0, r1), view) in enumerate(zip(self.chunks, self.views)):
                n = r1 - r0
                part1 = self.part1[n]
                mmv_partial(view, self.cols, self.v_rhs, self.sigma, part1, panel16=self.fwd[i])
                finish_w16(part1, self.T, self.Wf, self.absmax, self.W16, None if w is None else w[r0:r1])
                S = self.pslabs[i]
                panel16_tmm(self.fwd[i], self.W16, self.absmax, n, M, self.part3[slab:slab + S])
                slab += S
            self.have_fwd = True
            self.part1 = None                                       # only this pass needs the tile's slabs
            return finish_rows(self.part3, self.T, out, scale)
        if not self.have_tr:
            self._fill_transposed()
        # no kernel value is evaluated from here on: V -> fp16 split, then two panel passes per chunk
        self.Vpad[0, :, :self.T].copy_(v)
        finish_w16(self.Vpad, self.T, self.Vf, self.absmax_v, self.V16)
        slab = 0
        for i, (r0, r1) in enumerate(self.chunks):
            n = r1 - r0
            kv = self.kv_part[n]
            panel16_tmm(self.tr[i], self.V16, self.absmax_v, M, n, kv)                      # K_chunk v from K_chunk^T
            finish_w16(kv, self.T, self.Wf, self.absmax, self.W16, None if w is None else w[r0:r1])
            S = self.pslabs[i]
            panel16_tmm(self.fwd[i], self.W16, self.absmax, n, M, self.part3[slab:slab + S])  # K_chunk^T (K_chunk v + w)
            slab += S
        return finish_rows(self.part3, self.T, out, scale)


# EXPERIMENTAL precision tier: resident sweeps stream the hi plane only (K to 11 bits, 2 B per value) once the panels
# are filled.  ODF_PANEL_HI_ONLY=1.  Off by default: emulated on the CPU only so far (tools/precision_study.py).
PANEL_HI_ONLY = os.environ.get("ODF_PANEL_HI_ONLY", "0") not in ("0", "")
RESIDENT_NO_QUERY = 1 << 30  # panels up to this size are taken to fit without querying the free memory
RESIDENT_FRACTION = 0.85   # share of the free device memory the resident panels may take in mode "auto"
# keep only K_chunk (default) instead of K_chunk and K_chunk^T: K v then comes from the same panel through
# odf_panel16_mmv, half the memory and no transposed tile pass.  ODF_RESIDENT_SINGLE=0 selects the two-copy variant.
RESIDENT_SINGLE_COPY = os.environ.get("ODF_RESIDENT_SINGLE", "1") not in ("0", "")


def resident_bytes(n_rows, M):
    """Bytes of the resident fp16-plane panel sets (K and, unless RESIDENT_SINGLE_COPY, K^T) of an n_rows x M block."""
    L = _lib.load()
    chunk = _resident_chunk(n_rows)
    total = 0
    for r0 in range(0, n_rows, chunk):
        n = min(n_rows, r0 + chunk) - r0
        total += int(L.odf_panel16_bytes(n, M)) * (1 if RESIDENT_SINGLE_COPY else 2)
    return total


def _free_bytes(device):
    free, _total = torch.cuda.mem_get_info(device)
    # blocks cached by torch's allocator are reusable too
    return free + torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)


def resident_fits(n_rows, M, device):
    return resident_bytes(n_rows, M) <= RESIDENT_FRACTION * _free_bytes(device)


def resident_plan(n_rows, M, device, budget=None):
    """How many row chunks of an n_rows x M block can stay resident (mode "auto").  None = all of them: the panels fit
    into RESIDENT_FRACTION of the free device memory; otherwise (single-copy variant) as many as fit beside the
    transient panel the streamed chunks share; 0 = stream everything (mode "panel16")."""
    L = _lib.load()
    need = resident_bytes(n_rows, M)
    if budget is None and need <= RESIDENT_NO_QUERY:
        # small fits (the minibootstrap regime: a few MB of panel, hundreds of refits): cudaMemGetInfo costs ~3 ms on a
        # 180 GB device, a third of such a fit -- do not ask; an allocation failure falls back to streaming anyway
        return None
    budget = RESIDENT_FRACTION * _free_bytes(device) if budget is None else budget
    if need <= budget:
        return None                                                 # everything (chunks of _resident_chunk rows)
    if not RESIDENT_SINGLE_COPY:
        return 0
    chunk = _resident_chunk(n_rows, all_resident=False)
    per_chunk = int(L.odf_panel16_bytes(chunk, M))
    k = int((budget - per_chunk) // per_chunk)                      # one transient panel + k resident ones
    return max(0, min(k, n_rows // chunk))


# ---- RLS box refiners ---------------------------------------------------------------------------
def rls_train(X, Yw, perm, seg_host, row_class, lam):
    """All classes' ridge regressors in one call (odf_rls_train).  X [n x d] fp32 (cuda), Yw [n x 4] fp64 whitened targets
    per ORIGINAL row, perm [n_sel] int64 row indices sorted by class, seg_host: python list of n_classes + 1 boundaries in
    perm, row_class [n_sel] int32.  Returns (W [n_classes x 4 x (d + 1)] fp32, losses [n_sel x 4] fp32 in perm order)."""
    L = _lib.load()
    X = _req(X, "X", 2)
    X, ldx = _rowmajor(X)
    n_sel, d, C = int(perm.shape[0]), int(X.shape[1]), len(seg_host) - 1
    assert Yw.dtype == torch.float64 and Yw.is_contiguous() and Yw.shape == (X.shape[0], 4) and Yw.is_cuda
    assert perm.dtype == torch.int64 and perm.is_contiguous() and row_class.dtype == torch.int32 and row_class.is_contiguous()
    assert seg_host[0] == 0 and seg_host[-1] == n_sel
    dev = X.device
    W = torch.zeros((C, 4, d + 1), dtype=torch.float32, device=dev)
    losses = torch.empty((max(n_sel, 1), 4), dtype=torch.float32, device=dev)
    wsb = int(L.odf_rls_workspace_bytes(n_sel, d, C))
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    seg = (ctypes.c_int64 * (C + 1))(*[int(v) for v in seg_host])
    check(L.odf_rls_train(ptr(X), n_sel, d, ldx, ptr(Yw), ptr(perm), ctypes.cast(seg, ctypes.c_void_p), ptr(row_class), C,
                          float(lam), ptr(W), ptr(losses), ptr(ws), wsb, _stream()), "odf_rls_train")
    _count(4)
    return W, losses[:n_sel]


def rls_apply(feat, Wp, bias, Tinv, mu, ex_boxes, img_w, img_h, eps, mean=None, zscale=1.0):
    """Fused RLS apply + un-whitening + box decode for one image (odf_rls_apply): out [n x (C + 1) x 4]."""
    L = _lib.load()
    feat, ldf = _rowmajor(_req(feat, "feat", 2))
    n, d = int(feat.shape[0]), int(feat.shape[1])
    C = int(mu.shape[0])
    assert Wp.shape == (d, 4 * C) and Wp.is_contiguous() and bias.shape == (4 * C,) and Tinv.shape == (C, 4, 4) and Tinv.is_contiguous()
    ex = _req(ex_boxes, "boxes", 2).contiguous()
    out = torch.empty((n, C + 1, 4), dtype=torch.float32, device=feat.device)
    if n == 0:
        return out
    mean = None if mean is None else _req(mean, "mean").contiguous()
    check(L.odf_rls_apply(ptr(feat), n, d, ldf, ptr(Wp), ptr(bias.contiguous()), ptr(Tinv), ptr(mu.contiguous()), ptr(ex), C,
                          float(img_w), float(img_h), float(eps), ptr(mean), float(zscale), ptr(out), _stream()), "odf_rls_apply")
    _count(1)
    return out


# ---- index side: selection / gather (minibootstrap), box decode, detection post-processing ----
def select_indices(scores, thresh, strict=True):
    """Stable compaction: (idx, count) with idx[:count] == torch.where(scores > thresh)[0] (>= when not
    strict).  `count` stays on the device (int32 tensor); idx has len(scores) entries."""
    L = _lib.load()
    s = _req(scores, "scores")
    if s.dim() == 2 and s.shape[1] == 1:
        s = s[:, 0]
    assert s.dim() == 1
    n = int(s.shape[0])
    idx = torch.empty((max(n, 1),), dtype=torch.int64, device=s.device)
    count = torch.zeros((1,), dtype=torch.int32, device=s.device)
    wsb = int(L.odf_select_workspace_bytes(n))
    ws = torch.empty((max(wsb, 8),), dtype=torch.uint8, device=s.device)
    check(L.odf_select_indices(ptr(s), n, s.stride(0) if n else 1, float(thresh), 1 if strict else 0, ptr(idx),
                               ptr(count), ptr(ws), wsb, _stream()), "odf_select_indices")
    _count(3)
    return idx, count


def gather_rows(src, idx, count, dst, max_rows=None):
    """dst[k] = src[idx[k]] for k < count (device count)."""
    L = _lib.load()
    src = _req(src, "src", 2)
    assert src.stride(1) == 1 and dst.stride(1) == 1 and dst.shape[1] >= src.shape[1]
    mr = int(min(idx.shape[0], dst.shape[0]) if max_rows is None else max_rows)
    check(L.odf_gather_rows(ptr(src), src.stride(0), ptr(idx), ptr(count), mr, src.shape[1], ptr(dst), dst.stride(0),
                            _stream()), "odf_gather_rows")
    _count(1)
    return dst


def decode_boxes(ex_boxes, deltas, img_w, img_h):
    L = _lib.load()
    ex = _req(ex_boxes, "boxes", 2).contiguous()
    dl = _req(deltas, "deltas", 2).contiguous()
    R, Tc = int(dl.shape[0]), int(dl.shape[1]) // 4
    out = torch.empty_like(dl)
    check(L.odf_decode_boxes(ptr(ex), ptr(dl), R, Tc, float(img_w), float(img_h), ptr(out), _stream()), "odf_decode_boxes")
    _count(1)
    return out


def detect_postprocess(boxes, scores, score_thresh, nms_thresh, dets_per_img):
    """filter_results on the GPU.  Returns (boxes [k,4], scores [k], labels [k], rois [k]) trimmed to the
    surviving detections (one device->host read of the count)."""
    L = _lib.load()
    b = _req(boxes, "boxes", 2).contiguous()
    s = _req(scores, "scores", 2).contiguous()
    R, Tc = int(s.shape[0]), int(s.shape[1])
    assert b.shape[0] == R and b.shape[1] == 4 * Tc
    cap = max(R * (Tc - 1), 1)
    dev = s.device
    ob = torch.empty((cap, 4), dtype=torch.float32, device=dev)
    os_ = torch.empty((cap,), dtype=torch.float32, device=dev)
    ol = torch.empty((cap,), dtype=torch.int64, device=dev)
    orr = torch.empty((cap,), dtype=torch.int64, device=dev)
    cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
    wsb = int(L.odf_postprocess_workspace_bytes(R, Tc))
    ws = torch.empty((max(wsb, 8),), dtype=torch.uint8, device=dev)
    check(L.odf_detect_postprocess(ptr(b), ptr(s), R, Tc, float(score_thresh), float(nms_thresh), int(dets_per_img),
                                   ptr(ob), ptr(os_), ptr(ol), ptr(orr), ptr(cnt), ptr(ws), wsb, _stream()),
          "odf_detect_postprocess")
    _count(6)
    k = int(cnt.item())
    return ob[:k], os_[:k], ol[:k], orr[:k]
