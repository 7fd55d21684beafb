"""RLS box-refinement regressors, one 4-output ridge regressor per class.

Reference: src/modules/region-refiner/region_refiner_trainer/train_region_refiner.py:13-119
(same constructor/`__call__`/`train`/`solve` surface and the same model schema:
{'mu'(4), 'T'(4x4), 'T_inv'(4x4), 'Beta': {'0'..'3': {'weights'(d+1) fp32, 'losses'(n) fp32}}} in a
numpy object array, `None` entries for classes without samples).

Arithmetic is fp64 on the GPU, batched where the reference loops: the normal matrix [X 1]^T [X 1]
is one DSYRK-shaped product, one Cholesky and ONE triangular-solve pair for all four targets
(cuBLAS / cuSOLVER through torch.linalg — library-bound M^3-class work, like the FALKON
preconditioner).  `torch.eig` (removed from PyTorch) is replaced by `eigh`: S is a symmetric 4x4
matrix and T, T_inv are invariant to the order and sign of its eigenvectors.
"""
import os
import time

import numpy as np
import torch


class RegionRefinerTrainer:
    def __init__(self, cfg, lmbd, is_rpn):
        self.cfg = cfg
        self.lambd = lmbd
        self.COXY = None
        self.is_rpn = is_rpn

    def __call__(self, COXY, output_dir=None):
        self.COXY = COXY
        return self.train(output_dir=output_dir)

    def train(self, output_dir=None):
        classes = self.cfg["CHOSEN_CLASSES"]
        first = 0 if self.is_rpn else 1
        n_cls = len(classes)
        models = np.empty((0))
        dev = torch.device("cuda") if torch.cuda.is_available() else self.COXY["X"].device
        t0 = time.time()
        labels = self.COXY["C"]
        for i in range(first, n_cls):
            print("Training regressor for class %s (%d/%d)" % (classes[i], i, n_cls - 1))
            rows = torch.where(labels == i)[0]
            print("Training with %i examples" % len(rows))
            if len(rows) == 0:
                models = np.append(models, {"mu": None, "T": None, "T_inv": None, "Beta": None})
                print("No indices for class %s" % (classes[i]))
                continue
            Xi = self.COXY["X"][rows].to(device=dev, dtype=torch.float64)
            Yi = self.COXY["Y"][rows].to(device=dev, dtype=torch.float64)
            Xi = torch.cat((Xi, torch.ones((Xi.shape[0], 1), dtype=torch.float64, device=dev)), dim=1)
            # centre and whiten the 4-d targets
            mu = Yi.mean(dim=0)
            Yi = Yi - mu
            S = Yi.T @ Yi / Yi.shape[0]
            evals, W = torch.linalg.eigh(S)
            root = torch.sqrt(evals + 0.001)
            T = (W / root) @ W.T
            T_inv = (W * root) @ W.T
            Yi = Yi @ T
            Beta = self.solve(Xi, Yi, self.lambd)
            models = np.append(models, {"mu": mu.to(dev).float(), "T": T.to(dev).float(),
                                        "T_inv": T_inv.to(dev).float(), "Beta": Beta})
            mean_losses = torch.stack([Beta[k]["losses"].mean() for k in Beta])
            print("Mean losses:", mean_losses)
        training_time = time.time() - t0
        print("Time required to train %d regressors: %f seconds." % (n_cls - 1, training_time))
        if output_dir:
            who = "RPN's Online Region Refiner" if self.is_rpn else "Detector's Online Region Refiner"
            tail = " \n" if self.is_rpn else " \n \n"
            with open(os.path.join(output_dir, "result.txt"), "a") as fid:
                fid.write("{} training time: {}min:{}s{}".format(who, int(training_time / 60),
                                                                 round(training_time % 60), tail))
        return models

    def solve(self, X, y, lmbd, X_test=None, Y_test=None, indices=None):
        """w_k = (X^T X + lmbd I)^-1 X^T y_k for the four target columns; `indices` (optional,
        one row subset per target) reproduces the reference's per-target refits."""
        eye = torch.eye(X.shape[1], device=X.device, dtype=torch.float64)
        out = {}
        if indices is None:
            R = torch.linalg.cholesky(X.T @ X + lmbd * eye)
            Wk = torch.cholesky_solve(X.T @ y[:, :4], R)            # all four right-hand sides at once
            resid = X @ Wk - y[:, :4]
            for k in range(4):
                out[str(k)] = {"weights": Wk[:, k].contiguous().float(), "losses": (0.5 * resid[:, k] ** 2).float()}
            return out
        for k in range(4):
            Xk, yk = X[indices[k]], y[indices[k]][:, k]
            R = torch.linalg.cholesky(Xk.T @ Xk + lmbd * eye)
            w = torch.cholesky_solve((Xk.T @ yk)[:, None], R)[:, 0]
            out[str(k)] = {"weights": w.float(), "losses": (0.5 * (Xk @ w - yk) ** 2).float()}
        return out
