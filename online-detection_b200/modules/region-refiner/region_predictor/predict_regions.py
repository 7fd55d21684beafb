"""Apply the per-class RLS box refiners and decode the deltas (legacy +1 / eps box convention).

Reference: src/modules/region-refiner/region_predictor/predict_regions.py:7-80 — same surface
(`RegionPredictor(cfg, models)(boxes, features, normalize_features, stats)`), same output layout
(each BoxList's bbox becomes (n, num_classes, 4) with the un-refined box in slot 0).
All classes are applied by ONE fused libodf kernel per image (`odf_rls_apply`, csrc/odf_rls.cu): features @ [W_1 | W_2 | ...]
(d x 4(C-1)) + bias, the 4x4 un-whitening per class, the box decode and the clipping -- instead of the reference's
per-class loop of matmuls; the optional z-scoring of the features is fused into the kernel's feature load.
"""
import os
import sys

import numpy as np
import torch

_PKG = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), os.pardir, os.pardir, os.pardir))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)


class RegionPredictor:
    def __init__(self, cfg, models):
        self.cfg = cfg
        self.models = models
        self._packed = None

    def __call__(self, boxes, features, normalize_features=False, stats=None):
        # (the reference drops the two keyword arguments here; kept for parity, predict_regions.py:13)
        return self.predict(boxes, features, normalize_features=False, stats=None)

    def _pack(self, device):
        if self._packed is None or self._packed[0].device != device:
            Ws, bs, Tinv, mu = [], [], [], []
            for m in self.models:
                w = torch.stack([m["Beta"][str(k)]["weights"] for k in range(4)], dim=1).to(device)   # (d+1, 4)
                Ws.append(w[:-1])
                bs.append(w[-1])
                Tinv.append(m["T_inv"].to(device))
                mu.append(m["mu"].to(device))
            self._packed = (torch.cat(Ws, dim=1).float().contiguous(), torch.cat(bs).float().contiguous(),
                            torch.stack(Tinv).float().contiguous(), torch.stack(mu).float().contiguous())
        return self._packed

    def predict(self, boxes, features, normalize_features=False, stats=None):
        from odf import ops
        img_w, img_h = boxes[0].size
        dev = torch.device("cuda", torch.cuda.current_device())
        W, b, Tinv, mu = self._pack(dev)
        eps = float(np.spacing(1))
        mean, zscale = None, 1.0
        if normalize_features:
            mean = stats["mean"].to(dev).float()
            zscale = 20.0 / float(stats["mean_norm"].item())
        for i in range(len(boxes)):
            not_gt = np.nonzero(features[i]["gt"] == 0)
            feat = torch.as_tensor(features[i]["feat"][not_gt, :][0], dtype=torch.float32, device=dev)
            ex = boxes[i].bbox.to(dev).float()
            boxes[i].bbox = ops.rls_apply(feat, W, b, Tinv, mu, ex, img_w, img_h, eps, mean, zscale)
        return boxes
