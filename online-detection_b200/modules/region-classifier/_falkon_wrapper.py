"""Shared implementation of the two FALKONWrapper flavours.

Reference surface: src/modules/region-classifier/FALKONWrapper_with_centers_selection.py:16-95 and
FALKONWrapper_with_centers_selection_incore.py:16-99 — same constructor (YAML keys
[RPN.]ONLINE_REGION_CLASSIFIER.CLASSIFIER.{sigma,lambda,M} or ONLINE_SEGMENTATION.CLASSIFIER.*,
defaults sigma=5, lambda=1e-3), train() returning a deep copy of the fitted model, predict() as
model.predict(X), and the <= M/2-positives centre-selection rule with replacement sampling.
The model is odf.InCoreFalkon / odf.Falkon: hand-written sm_100a kernels, no `falkon` package.
"""
import copy
import os
import sys

import torch
import yaml

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import _paths  # noqa: E402,F401
import ClassifierAbstract as ca  # noqa: E402
from MyCenterSelector import MyCenterSelector  # noqa: E402
from odf import Falkon, FalkonOptions, GaussianKernel, InCoreFalkon  # noqa: E402

kernels = type("kernels", (), {"GaussianKernel": GaussianKernel})   # `kernels.GaussianKernel(sigma=...)`


def _classifier_options(cfg, is_rpn, is_segmentation):
    if is_rpn:
        cfg = cfg["RPN"]
    section = "ONLINE_SEGMENTATION" if is_segmentation else "ONLINE_REGION_CLASSIFIER"
    return cfg, cfg[section]["CLASSIFIER"]


class FALKONWrapperBase(ca.ClassifierAbstract):
    MODEL_CLS = InCoreFalkon
    IN_CORE = True

    def __init__(self, cfg_path=None, is_rpn=False, is_segmentation=False):
        if cfg_path is not None:
            with open(cfg_path) as fh:
                self.cfg = yaml.load(fh, Loader=yaml.FullLoader)
        self.cfg, opts = _classifier_options(self.cfg, is_rpn, is_segmentation)
        if "sigma" in opts:
            self.sigma = opts["sigma"]
        else:
            print("Sigma not given for creating Falkon, default value is used.")
            self.sigma = 5
        if "lambda" in opts:
            self.lam = opts["lambda"]
        else:
            print("Lambda not given for creating Falkon, default value is used.")
            self.lam = 0.001
        self.kernel = None
        self.nyst_centers = opts["M"]
        self.maxiter = 20       # FALKON's default number of CG iterations (…incore.py:41)
        self.model = None

    # -- a3: centre indices: at most M/2 positives, negatives fill up, both WITH replacement
    def compute_indices_selection(self, y):
        budget = self.nyst_centers
        half = int(budget / 2)
        pos = (y == 1).nonzero()
        if pos.size(0) > half:
            pos = pos[torch.randint(pos.size(0), (half,))]
        neg = (y == -1).nonzero()
        room = budget - pos.size(0)
        if neg.size(0) > room:
            neg = neg[torch.randint(neg.size(0), (room,))]
        return torch.cat((pos, neg), dim=0).squeeze().tolist()

    def _options(self):
        if self.IN_CORE:
            return FalkonOptions(min_cuda_iter_size_32=0, min_cuda_iter_size_64=0, keops_active="no",
                                 min_cuda_pc_size_32=0, min_cuda_pc_size_64=0, store_kernel_d_threshold=250)
        return FalkonOptions(min_cuda_iter_size_32=0, min_cuda_iter_size_64=0, keops_active="no")

    def train(self, X, y, sigma=None, lam=None):
        sigma = self.sigma if sigma is None else sigma
        lam = self.lam if lam is None else lam
        self.kernel = kernels.GaussianKernel(sigma=sigma)
        indices = self.compute_indices_selection(y)
        if isinstance(indices, int):
            indices = [indices]
        selector = MyCenterSelector(indices)
        self.model = self.MODEL_CLS(kernel=self.kernel, penalty=lam, M=len(indices), maxiter=self.maxiter,
                                    center_selection=selector, options=self._options())
        if self.model is None:
            print("Model is None in trainRegionClassifier function")
            sys.exit(0)
        src_device = X.device
        if not self.IN_CORE and not X.is_cuda:
            # out-of-core flavour: features are parked in host RAM, the iterations still run on
            # the GPU (min_cuda_iter_size_* = 0 upstream) — fit() uploads them on a side stream while
            # the centres are prepared and the preconditioner is built
            self.model.fit(X, y)
            self.model.ny_points_ = self.model.ny_points_.to(src_device)
            self.model.alpha_ = self.model.alpha_.to(src_device)
        else:
            self.model.fit(X, y)
        return copy.deepcopy(self.model)

    def predict(self, model, X_np, y=None):
        if X_np.is_cuda:
            if not model.ny_points_.is_cuda:
                model.ny_points_ = model.ny_points_.to(X_np.device)
                model.alpha_ = model.alpha_.to(X_np.device)
            return model.predict(X_np) if y is None else model.predict(X_np, y)
        # host-resident features (out-of-core flavour): score on the GPU, hand back on the host
        dev = torch.device("cuda")
        if not model.ny_points_.is_cuda:
            staged = copy.copy(model)
            staged.ny_points_ = model.ny_points_.to(dev)
            staged.alpha_ = model.alpha_.to(dev)
        else:
            staged = model
        return staged.predict(X_np.to(dev)).to(X_np.device)

    def test(self):
        pass
