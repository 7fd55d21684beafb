"""CPU tests of the host logic of the EXPERIMENTAL paths that have not run on a GPU yet: the routing of
ops.gemm onto the NT-only split GEMM (explicit transposes for the other three forms) and the k-slicing of
ops.gemm_nt_split, with the device calls replaced by torch emulations."""
import pytest
import torch


@pytest.fixture()
def split_emu(monkeypatch, lib):
    from odf import ops
    calls = []

    def fake_nt(A, B, C, alpha=1.0, beta=0.0, kslice=None, kind=None):
        assert A.shape[1] == B.shape[1] and C.shape == (A.shape[0], B.shape[0])
        assert A.stride(1) == 1 and B.stride(1) == 1
        calls.append((tuple(A.shape), tuple(B.shape)))
        C.copy_(alpha * (A.double() @ B.double().T).float() + (beta * C if beta != 0.0 else 0.0))
        return C

    monkeypatch.setattr(ops, "gemm_nt_split", fake_nt)
    monkeypatch.setattr(ops, "GEMM_SPLIT", True)
    monkeypatch.setattr(ops, "GEMM_SPLIT_MIN", 8)
    return ops, calls


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_routes_every_form_onto_the_nt_kernel(split_emu, ta, tb):
    ops, calls = split_emu
    g = torch.Generator().manual_seed(1)
    m, n, k = 24, 17, 33
    A = torch.randn((k, m) if ta else (m, k), generator=g)
    B = torch.randn((n, k) if tb else (k, n), generator=g)
    C = torch.randn(m, n, generator=g)
    ref = -0.5 * ((A.T if ta else A).double() @ (B.T if tb else B).double()) + 2.0 * C.double()
    ops.gemm(A, B, C, trans_a=ta, trans_b=tb, alpha=-0.5, beta=2.0)
    assert calls == [((m, k), (n, k))]
    assert float((C.double() - ref).abs().max()) < 1e-5


def test_blocked_build_uses_only_gemm_for_the_cubic_work(split_emu):
    """With the split GEMM switched on every large product of the blocked preconditioner build goes through it."""
    import cpu_backend as be
    from odf import precond_blocked as pb
    ops, calls = split_emu

    class Table:                                   # the CPU table with ops.gemm (-> fake split kernel) plugged in
        potrf_upper_ = staticmethod(be.potrf_upper_)
        precond_solve_ = staticmethod(be.precond_solve_)
        add_diag_ = staticmethod(be.add_diag_)
        zero_lower_ = staticmethod(be.zero_lower_)
        axpby = staticmethod(be.axpby)
        gemm = staticmethod(ops.gemm)

    g = torch.Generator().manual_seed(0)
    M = 96
    R = torch.randn(M, M, generator=g)
    K = (R @ R.T / M + 0.5 * torch.eye(M)).float()
    Tm, Am = pb.build(Table, K.clone(), 1e-3, 1e-5, nb=32)
    T0 = torch.linalg.cholesky(K.double() + 1e-5 * M * torch.eye(M, dtype=torch.float64), upper=True)
    A0 = torch.linalg.cholesky(T0 @ T0.T / M + 1e-3 * torch.eye(M, dtype=torch.float64), upper=True)
    assert float((Tm.double() - T0).abs().max()) < 1e-4 and float((Am.double() - A0).abs().max()) < 1e-4
    assert len(calls) == 2 * (2 + 1) + 3 + 0       # two factorisations x (3 trailing updates) + 3 block columns of T T^T
