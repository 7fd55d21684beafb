"""Blocked build of the FALKON preconditioner on top of the operator table (`odf.ops` on the GPU, the test-only CPU
table in tests/cpu_backend.py), so that every O(M^3) flop of it runs through ONE primitive, `be.gemm`:

    T = chol_upper(K_MM + eps M I)          blocked right-looking Cholesky: nb x nb diagonal factorisations
    G = T T^T / M + lam I                   (be.potrf_upper_), row-panel solves (be.precond_solve_), trailing
    A = chol_upper(G)                       updates and the triangle-aware T T^T as GEMMs

(falkon `FalkonPreconditioner.init`; reached from InCoreFalkon.fit,
src/modules/region-classifier/FALKONWrapper_with_centers_selection_incore.py:68.)

Why it exists: `odf_precond_init` hands the factorisations to cuSOLVER potrf (SIMT fp32, 32 TFLOP/s at M = 10 k) and the
products to cuBLAS sgemm (67 TFLOP/s); with resident sweeps that build is 18 % of the single-GPU fit and more than half
of the 8-GPU one.  Here the factorisation is expressed as GEMMs with a long inner dimension, which is
the shape the 3-pass split-fp16 tensor-core GEMM planned for round 2 is good at; swapping that kernel in is then a change
of `be.gemm` only.  EXPERIMENTAL and off by default (`FalkonOptions(precond_build="blocked")`): with today's `be.gemm`
(cuBLAS sgemm) it does the same flops as the library path; the host logic is pinned on the CPU against
torch.linalg.cholesky (tests/test_host_logic.py) and behind ODF_EXPERIMENTAL=1 on the GPU.
"""
import torch

from ._lib import ODF_SOLVE_TT as SOLVE_TT

DEFAULT_BLOCK = 1024


def potrf_upper_blocked_(be, A, nb=DEFAULT_BLOCK):
    """In-place upper Cholesky factor of the symmetric positive definite matrix whose UPPER triangle is in A
    (A = U^T U; the strict lower triangle is neither read nor written).  Right-looking, block size nb:

        for each diagonal block j:   U_jj = chol(A_jj)                        be.potrf_upper_ on a contiguous copy
                                     U_j,> = U_jj^-T A_j,>                     be.precond_solve_ (TRSM, nb x nb factor)
                                     A_>,> -= U_j,>^T U_j,>   (upper blocks)   be.gemm, one call per block column

    The trailing update of block column J touches rows j1..J1 only (the upper triangle): M^3/3 flops in total."""
    M = A.shape[0]
    assert A.dim() == 2 and A.shape[1] == M and A.stride(1) == 1
    for j0 in range(0, M, nb):
        j1 = min(M, j0 + nb)
        D = A[j0:j1, j0:j1].contiguous()
        be.potrf_upper_(D)
        A[j0:j1, j0:j1].copy_(D)
        if j1 == M:
            break
        P = A[j0:j1, j1:]                                    # row panel, overwritten with U_j,>
        be.precond_solve_(D, P, SOLVE_TT)
        for J0 in range(j1, M, nb):
            J1 = min(M, J0 + nb)
            # A[j1:J1, J0:J1] -= U[j, j1:J1]^T U[j, J0:J1]
            be.gemm(A[j0:j1, j1:J1], A[j0:j1, J0:J1], A[j1:J1, J0:J1], trans_a=True, alpha=-1.0, beta=1.0)
    return A


def ttt_upper(be, T, out=None, nb=DEFAULT_BLOCK):
    """Upper triangle (block-wise) of T T^T for an upper-triangular T: block column J needs only
    T[0:J1, J0:] . T[J0:J1, J0:]^T because T[J0:J1, k] = 0 for k < J0 -- M^3/3 multiply-adds instead of M^3.
    Entries below the block diagonal of `out` are left as they are (potrf never reads them)."""
    M = T.shape[0]
    if out is None:
        out = torch.zeros((M, M), dtype=T.dtype, device=T.device)
    for J0 in range(0, M, nb):
        J1 = min(M, J0 + nb)
        be.gemm(T[0:J1, J0:], T[J0:J1, J0:], out[0:J1, J0:J1], trans_b=True)
    return out


def build(be, Kmm, lam, eps, nb=DEFAULT_BLOCK):
    """(T, A) from K_MM (overwritten), as odf_precond_init returns them: both upper triangular with a ZERO strict
    lower triangle (the explicit-inverse and TRSM applications read them as plain matrices)."""
    M = Kmm.shape[0]
    be.add_diag_(Kmm, eps * M)
    Tm = potrf_upper_blocked_(be, Kmm, nb)
    be.zero_lower_(Tm)
    G = ttt_upper(be, Tm, nb=nb)
    be.axpby(G, 1.0 / M, G)
    be.add_diag_(G, lam)
    Am = potrf_upper_blocked_(be, G, nb)
    be.zero_lower_(Am)
    return Tm, Am
