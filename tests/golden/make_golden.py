"""Regenerates the committed golden fixtures from the CPU oracle (fp64).

The third-party `falkon` arithmetic of the reference cannot run in this container (falkon is not
installed), so THESE vectors pin the ORACLE against regressions and give the GPU
tests fixed inputs; they are not outputs of the reference itself.  (Everything around that
arithmetic is pinned by tests/golden/make_reference_golden.py, which runs the reference's own
first-party modules.)

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import falkon_oracle as orc  # noqa: E402


def main():
    N, d, T, M, sigma, lam = 900, 48, 3, 96, 12.0, 1e-4
    X, c, Y = orc.make_synthetic(N, d, T, seed=0)
    idx = orc.shared_centres(c, M, seed=1)
    C = X[idx]
    # fp64 arithmetic with the fp32 epsilons the GPU path uses (pc 1e-5, cg 1e-7)
    alpha = orc.falkon_fit(X, Y, C, sigma, lam, dtype=torch.float64, eps_pc=1e-5, eps_cg=1e-7)
    scores = orc.falkon_predict(X[:64], C, alpha, sigma)
    np.savez_compressed(os.path.join(HERE, "falkon_small.npz"), X=X.numpy(), Y=Y.numpy(), centre_idx=idx.numpy(),
                        sigma=sigma, lam=lam, alpha=alpha.numpy(), scores=scores.numpy())

    g = torch.Generator().manual_seed(5)
    n, dd = 400, 24
    Xr = torch.randn(n, dd, generator=g)
    Wt = torch.randn(dd, 4, generator=g) * 0.1
    Yr = Xr @ Wt + 0.05 * torch.randn(n, 4, generator=g) + torch.tensor([0.1, -0.2, 0.05, 0.0])
    m = orc.rls_train_class(Xr, Yr, 10.0)
    W = torch.stack([m["Beta"][str(k)]["weights"] for k in range(4)], 1)
    np.savez_compressed(os.path.join(HERE, "rls_small.npz"), X=Xr.numpy(), Y=Yr.numpy(), lam=10.0, W=W.numpy(),
                        mu=m["mu"].numpy(), T=m["T"].numpy(), T_inv=m["T_inv"].numpy())
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
