"""TEST-ONLY stand-in for odf.ops with the same names, doing the arithmetic with the CPU oracle.
It exists so that the row-sharding / all-reduce / CG host logic of odf.Falkon can be exercised
with world_size 2 over gloo in a container without a GPU.  Never used by the product path."""
import torch

from oracle import falkon_oracle as orc

DT = torch.float64


class Prepared:
    def __init__(self, X, mean=None, scale=1.0, kind=None):
        Xp = X.to(DT)
        if mean is not None:
            Xp = Xp - mean.to(DT)
        self.X = Xp * scale
        self.n, self.d = X.shape
        self.kind = 0
        self.hi = X          # only .device is consulted


def kmm(prep, sigma, out=None):
    return orc.gaussian_kernel(prep.X, prep.X, sigma, DT)


def precond_init(K, lam, eps):
    M = K.shape[0]
    T = torch.linalg.cholesky(K + eps * M * torch.eye(M, dtype=DT), upper=True)
    A = torch.linalg.cholesky(T @ T.T / M + lam * torch.eye(M, dtype=DT), upper=True)
    return T, A


def potrf_upper_(A):
    A.copy_(torch.linalg.cholesky(A, upper=True))
    return A


def add_diag_(A, value):
    A.diagonal().add_(value)
    return A


def zero_lower_(A):
    A.triu_()
    return A


def precond_solve_(Tri, B, which):
    tr = which in (1, 3)
    sol = torch.linalg.solve_triangular(Tri.T if tr else Tri, B.to(DT), upper=not tr)
    B.copy_(sol.to(B.dtype))
    return B


def precond_invert(Tri):
    return torch.linalg.solve_triangular(Tri, torch.eye(Tri.shape[0], dtype=Tri.dtype), upper=True)


def precond_apply(Inv, Bin, Bout, transposed):
    Bout.copy_(((Inv.T if transposed else Inv) @ Bin.to(Inv.dtype)).to(Bout.dtype))
    return Bout


def gemm(A, B, C, trans_a=False, trans_b=False, alpha=1.0, beta=0.0):
    res = alpha * ((A.T if trans_a else A) @ (B.T if trans_b else B))
    C.copy_(res + beta * C if beta != 0.0 else res)
    return C


def precond_apply_rows(Inv, r0, r1, Bin, Bout_rows, transposed):
    full = (Inv.T if transposed else Inv) @ Bin.to(Inv.dtype)
    Bout_rows[:r1 - r0].copy_(full[r0:r1].to(Bout_rows.dtype))
    return Bout_rows


class Sweeper:
    def __init__(self, rows, cols, sigma, T, mode="panel"):
        self.rows, self.cols, self.sigma, self.T = rows, cols, sigma, T

    def dmmv(self, v, w, out, scale=1.0, w_scale=1.0):
        res = orc.dmmv(self.rows.X, self.cols.X, None if v is None else v.to(DT),
                       None if w is None else w.to(DT) * w_scale, self.sigma, DT)
        out.copy_((res * scale).to(out.dtype))
        return out


def mmv_into(rows, cols, v, sigma, out):
    out.copy_(orc.mmv(rows.X, cols.X, v.to(DT), sigma, DT).to(out.dtype))


def axpby(out, alpha, A, beta=0.0, B=None):
    res = alpha * A
    if B is not None:
        res = res + beta * B
    out.copy_(res)
    return out


class CgState:
    def __init__(self, M, T, device):
        self.T = T
        self.state = torch.zeros(4 * T + 4)

    @property
    def converged_flag(self):
        return self.state[4 * self.T:4 * self.T + 1]

    def _frozen(self):
        return float(self.state[4 * self.T]) != 0.0

    def init(self, R):
        T = self.T
        self.state[:T] = (R.double() ** 2).sum(0).float()
        self.state[4 * T] = 0.0

    def alpha(self, P, AP, eps):
        T = self.T
        pap = (P.double() * AP.double()).sum(0).float()
        self.state[T:2 * T] = 0.0 if self._frozen() else self.state[:T] / (pap + eps)

    def axpy_a(self, Y, X, sign):
        if not self._frozen():
            Y += sign * X * self.state[self.T:2 * self.T]

    def residual(self, R, Bm, H):
        if not self._frozen():
            R.copy_(Bm - H)

    def beta(self, R, eps, tol):
        T = self.T
        if self._frozen():
            self.state[2 * T:3 * T] = 0.0
            return
        rs_new = (R.double() ** 2).sum(0).float()
        if float(rs_new.abs().max()) ** 0.5 < tol:
            self.state[2 * T:3 * T] = 0.0
            self.state[4 * T] = 1.0
            return
        self.state[2 * T:3 * T] = rs_new / (self.state[:T] + eps)
        self.state[:T] = rs_new

    def xpby_b(self, P, R):
        if not self._frozen():
            P.copy_(R + P * self.state[2 * self.T:3 * self.T])


def zscore_(X, mean, scale):
    if mean is not None:
        X -= mean
    X *= scale
    return X
