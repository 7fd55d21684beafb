"""Round-2 sizing probe for the tensor-core preconditioner (one GPU): what do the building blocks cost?

    potrf (cuSOLVER) and trsm-vs-identity (cuBLAS) at the block sizes a blocked Cholesky would use,
    cuBLAS sgemm vs the 3-pass split GEMM (odf_gemm_nt_split, operand pre-pass timed separately) at the shapes of the
    panel solve / trailing update / T T^T, and the accuracy of the split GEMM against fp64.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "online-detection_b200"))
from odf import ops, _lib  # noqa: E402
from odf._lib import check, ptr  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def spd(n, gen):
    A = torch.randn(n, n + 64, device="cuda", generator=gen)
    return A @ A.T / n + 0.1 * torch.eye(n, device="cuda")


def main():
    L = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(0)
    print("== cuSOLVER potrf / cuBLAS trsm (identity rhs) ==", flush=True)
    for n in (128, 256, 512, 1024, 2048, 4096):
        A = spd(n, g)
        buf = A.clone()

        def f_potrf():
            buf.copy_(A)
            ops.potrf_upper_(buf)
        t_copy = timed(lambda: buf.copy_(A))
        t_p = timed(f_potrf) - t_copy
        U = buf.clone()
        E = torch.eye(n, device="cuda")
        X = E.clone()

        def f_trsm():
            X.copy_(E)
            ops.precond_solve_(U, X, _lib.ODF_SOLVE_T)
        t_t = timed(f_trsm) - t_copy
        t_inv = timed(lambda: ops.precond_invert(U))
        print("n=%5d  potrf %.3f ms (%.1f TF/s)   trsm(I) %.3f ms   precond_invert %.3f ms" %
              (n, t_p, n ** 3 / 3 / t_p / 1e9, t_t, t_inv), flush=True)

    print("== sgemm vs split GEMM (NT) ==", flush=True)
    shapes = [(1024, 1024, 1024), (2048, 2048, 1024), (4096, 1024, 1024), (9000, 1024, 1024), (4096, 4096, 1024),
              (9000, 9000, 1024), (9000, 2048, 1024), (512, 512, 512), (256, 256, 256), (10000, 2048, 2048)]
    for (m, n, k) in shapes:
        A = torch.randn(m, k, device="cuda", generator=g)
        B = torch.randn(n, k, device="cuda", generator=g)
        C = torch.zeros(m, n, device="cuda")
        t_s = timed(lambda: ops.gemm(A, B, C, trans_b=True))
        C_s = C.clone()
        t_all = timed(lambda: ops.gemm_nt_split(A, B, C))
        pa = ops.Prepared(A, linear=True)
        pb = ops.Prepared(B, kind=pa.kind, linear=True)
        t_prep = timed(lambda: (ops.Prepared(A, linear=True), ops.Prepared(B, kind=pa.kind, linear=True)))

        def f_k():
            check(L.odf_gemm_nt_split(pa.kind, ptr(pa.hi), ptr(pa.lo), ptr(pa.sqn), ptr(pa.opscale), m, ptr(pb.hi), ptr(pb.lo),
                                      ptr(pb.sqn), ptr(pb.opscale), n, k, 1.0, 0.0, ptr(C), C.stride(0), ops._stream()), "gemm")
        t_k = timed(f_k)
        rows = slice(0, min(m, 512))
        ref = A[rows].double() @ B.double().T
        scale = A[rows].double().abs() @ B.double().abs().T
        e_split = float(((C[rows].double() - ref).abs() / scale).max())
        e_sg = float(((C_s[rows].double() - ref).abs() / scale).max())
        fl = 2.0 * m * n * k
        print("m=%5d n=%5d k=%5d  sgemm %.3f ms (%.0f TF/s, err %.1e)   split: kernel %.3f ms (%.0f TF/s)  prep %.3f ms  "
              "whole call %.3f ms  err %.1e" % (m, n, k, t_s, fl / t_s / 1e9, e_sg, t_k, fl / t_k / 1e9, t_prep, t_all, e_split), flush=True)


if __name__ == "__main__":
    main()
