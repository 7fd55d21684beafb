"""RLS box-refinement regressors, one 4-output ridge regressor per class.

Reference: src/modules/region-refiner/region_refiner_trainer/train_region_refiner.py:13-119
(same constructor/`__call__`/`train`/`solve` surface and the same model schema:
{'mu'(4), 'T'(4x4), 'T_inv'(4x4), 'Beta': {'0'..'3': {'weights'(d+1) fp32, 'losses'(n) fp32}}} in a
numpy object array, `None` entries for classes without samples).

Arithmetic is fp64 as in the reference, and ALL classes are trained by one call into libodf (`odf_rls_train`,
csrc/odf_rls.cu) where the reference loops over them: the rows are sorted by class once, the normal matrices
[X 1 y']^T [X 1 y'] of every class come from one launch on the fp64 tensor cores (the features stay fp32 in HBM and
are widened on the fly), the factorisations / solves are cuSOLVER calls on side streams, the losses one more pass.
The 4 x 4 target statistics (mean, covariance, whitening T / T_inv) are a few batched torch ops on (classes x 4 x 4)
tensors.  `torch.eig` (removed from PyTorch) is replaced by `eigh`: S is a symmetric 4x4 matrix and T, T_inv are
invariant to the order and sign of its eigenvectors.  No CPU path: without a CUDA device `train` raises.
"""
import os
import sys
import time

import numpy as np
import torch

_PKG = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), os.pardir, os.pardir, os.pardir))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)


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
        from odf import ops
        if not torch.cuda.is_available():
            raise RuntimeError("RegionRefinerTrainer needs a CUDA device (libodf has no CPU fallback)")
        classes = self.cfg["CHOSEN_CLASSES"]
        first = 0 if self.is_rpn else 1
        n_cls = len(classes)
        C = n_cls - first
        dev = torch.device("cuda", torch.cuda.current_device())
        t0 = time.time()
        lab = self.COXY["C"].reshape(-1).to(dev).long()
        X = self.COXY["X"].to(device=dev, dtype=torch.float32)
        Y = self.COXY["Y"].to(device=dev, dtype=torch.float64)
        # rows of the trained classes, sorted by class (stable: ascending row index inside a class, as `X[rows]` upstream)
        cls_of = lab - first
        sel = ((cls_of >= 0) & (cls_of < C)).nonzero()[:, 0]
        order = torch.sort(cls_of[sel], stable=True).indices
        perm = sel[order].contiguous()
        row_class = cls_of[perm].to(torch.int32).contiguous()
        counts = torch.bincount(row_class.long(), minlength=C)
        counts_h = counts.cpu().tolist()
        seg = [0]
        for n_c in counts_h:
            seg.append(seg[-1] + int(n_c))
        # target statistics per class, batched: mu, S = Yc^T Yc / n, T = W (D + 1e-3)^-1/2 W^T, T_inv = W (D + 1e-3)^1/2 W^T
        rc = row_class.long()
        cnt = counts.clamp(min=1).to(torch.float64)
        mu = torch.zeros((C, 4), dtype=torch.float64, device=dev).index_add_(0, rc, Y[perm]) / cnt[:, None]
        Yc = Y[perm] - mu[rc]
        S = torch.zeros((C, 16), dtype=torch.float64, device=dev).index_add_(
            0, rc, (Yc[:, :, None] * Yc[:, None, :]).reshape(-1, 16)).view(C, 4, 4) / cnt[:, None, None]
        evals, Wv = torch.linalg.eigh(S)
        root = torch.sqrt(evals + 0.001)
        T = (Wv / root[:, None, :]) @ Wv.transpose(1, 2)
        T_inv = (Wv * root[:, None, :]) @ Wv.transpose(1, 2)
        Yw = torch.zeros((X.shape[0], 4), dtype=torch.float64, device=dev)
        Yw[perm] = torch.einsum("ni,nij->nj", Yc, T[rc])
        if perm.numel() > 0:
            Wts, losses = ops.rls_train(X, Yw, perm, seg, row_class, float(self.lambd))
        models = np.empty((0))
        for c in range(C):
            i = c + first
            print("Training regressor for class %s (%d/%d)" % (classes[i], i, n_cls - 1))
            print("Training with %i examples" % counts_h[c])
            if counts_h[c] == 0:
                models = np.append(models, {"mu": None, "T": None, "T_inv": None, "Beta": None})
                print("No indices for class %s" % (classes[i]))
                continue
            Beta = {str(k): {"weights": Wts[c, k].contiguous(), "losses": losses[seg[c]:seg[c + 1], k].contiguous()} for k in range(4)}
            models = np.append(models, {"mu": mu[c].float(), "T": T[c].float(), "T_inv": T_inv[c].float(), "Beta": Beta})
            print("Mean losses:", losses[seg[c]:seg[c + 1]].mean(0))
        torch.cuda.synchronize(dev)
        training_time = time.time() - t0
        self.train_seconds_ = training_time
        print("Time required to train %d regressors: %f seconds." % (n_cls - 1, training_time))
        if output_dir:
            who = "RPN's Online Region Refiner" if self.is_rpn else "Detector's Online Region Refiner"
            tail = " \n" if self.is_rpn else " \n \n"
            with open(os.path.join(output_dir, "result.txt"), "a") as fid:
                fid.write("{} training time: {}min:{}s{}".format(who, int(training_time / 60),
                                                                 round(training_time % 60), tail))
        return models

    def solve(self, X, y, lmbd, X_test=None, Y_test=None, indices=None):
        """w_k = (X^T X + lmbd I)^-1 X^T y_k for the four target columns; `indices` (optional, one row subset per target)
        reproduces the reference's per-target refits.  Kept for callers that use the reference's `solve` directly (fp64
        torch.linalg on the tensors' device); `train` does not come through here -- it batches all classes in libodf."""
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
