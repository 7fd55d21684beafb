"""Apply the per-class RLS box refiners and decode the deltas (legacy +1 / eps box convention).

Reference: src/modules/region-refiner/region_predictor/predict_regions.py:7-80 — same surface
(`RegionPredictor(cfg, models)(boxes, features, normalize_features, stats)`), same output layout
(each BoxList's bbox becomes (n, num_classes, 4) with the un-refined box in slot 0).
All classes are applied in ONE GEMM: features @ [W_1 | W_2 | ...] (d x 4(C-1)), followed by a
batched 4x4 un-whitening, instead of the reference's per-class loop.
"""
import numpy as np
import torch


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
            self._packed = (torch.cat(Ws, dim=1), torch.cat(bs), torch.stack(Tinv), torch.stack(mu))
        return self._packed

    def predict(self, boxes, features, normalize_features=False, stats=None):
        n_cls = len(self.cfg["CHOSEN_CLASSES"])
        img_w, img_h = boxes[0].size
        dev = torch.device("cuda")
        W, b, Tinv, mu = self._pack(dev)
        eps = float(np.spacing(1))
        for i in range(len(boxes)):
            not_gt = np.nonzero(features[i]["gt"] == 0)
            feat = torch.tensor(features[i]["feat"][not_gt, :][0], device=dev)
            if normalize_features:
                feat = (feat - stats["mean"]) * (20 / stats["mean_norm"].item())
            ex = boxes[i].bbox.to(dev)
            n = ex.shape[0]
            Y = (feat @ W + b).view(n, n_cls - 1, 4)
            Y = torch.einsum("nck,ckj->ncj", Y, Tinv) + mu                    # un-whiten per class
            src_w = (ex[:, 2] - ex[:, 0] + eps)[:, None]
            src_h = (ex[:, 3] - ex[:, 1] + eps)[:, None]
            ctr_x = ex[:, 0:1] + 0.5 * src_w
            ctr_y = ex[:, 1:2] + 0.5 * src_h
            pcx = Y[..., 0] * src_w + ctr_x
            pcy = Y[..., 1] * src_h + ctr_y
            pw = torch.exp(Y[..., 2]) * src_w
            ph = torch.exp(Y[..., 3]) * src_h
            pred = torch.stack(((pcx - 0.5 * pw).clamp(min=0), (pcy - 0.5 * ph).clamp(min=0),
                                (pcx + 0.5 * pw - 1).clamp(max=img_w - 1), (pcy + 0.5 * ph - 1).clamp(max=img_h - 1)),
                               dim=2)
            boxes[i].bbox = torch.cat((ex[:, None, :], pred), dim=1)
        return boxes
