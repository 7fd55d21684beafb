"""Golden vectors for the detection post-processor: the REFERENCE's own `OnlineDetectionPostProcessor.forward` /
`filter_results` (src/modules/accuracy-evaluator/OnlineDetectionPostProcessor.py:12-79) and its first-party IoU twin
`compute_overlap_torch` (mrcnn_modified/utils/evaluations.py:4-18), run on synthetic proposals.

First-party and run as is: the box decode (py_od_utils.decode_boxes_detector), the strict `>` score threshold, the
per-class loop and concatenation order, the labels, the kthvalue top-K with `>=` (ties may keep more than K).
Third-party (maskrcnn-benchmark, un-vendored) and therefore shimmed with SURVEY Appendix B's conventions: the
`PostProcessor` base (`prepare_boxlist`), `BoxList` (resize to the same size, clip_to_image with TO_REMOVE = 1, indexing),
`cat_boxlist`, and `boxlist_nms` (= the oracle's `nms_plus1`: greedy, IoU(+1) > thresh, keep indices ascending).

Output: tests/golden/reference_post.npz, checked against the oracle's decode + filter_results by
tests/test_reference_golden.py (the CUDA post-processor is checked against the oracle in tests/test_gpu_post.py).

    python tests/golden/make_reference_golden_post.py      # needs /root/reference
"""
import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as base  # noqa: E402
from oracle import falkon_oracle as orc  # noqa: E402


class BoxList:
    def __init__(self, bbox, image_size, mode="xyxy"):
        self.bbox, self.size, self.mode, self.extra_fields = torch.as_tensor(bbox), image_size, mode, {}

    def add_field(self, k, v):
        self.extra_fields[k] = v

    def get_field(self, k):
        return self.extra_fields[k]

    def fields(self):
        return list(self.extra_fields)

    def resize(self, size):
        assert tuple(size) == tuple(self.size)      # the experiments resize to the image's own size
        return self

    def clip_to_image(self, remove_empty=True):
        assert not remove_empty
        w, h = self.size
        self.bbox[:, 0::2].clamp_(min=0, max=w - 1)
        self.bbox[:, 1::2].clamp_(min=0, max=h - 1)
        return self

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]


def boxlist_nms(boxlist, nms_thresh, max_proposals=-1, score_field="scores"):
    if nms_thresh <= 0:
        return boxlist
    keep = orc.nms_plus1(boxlist.bbox.numpy(), boxlist.get_field(score_field).numpy(), nms_thresh)
    return boxlist[torch.as_tensor(np.asarray(keep), dtype=torch.int64)]


def cat_boxlist(bboxes):
    out = BoxList(torch.cat([b.bbox for b in bboxes], 0), bboxes[0].size, bboxes[0].mode)
    for f in bboxes[0].fields():
        out.add_field(f, torch.cat([b.get_field(f) for b in bboxes], 0))
    return out


class PostProcessor(torch.nn.Module):
    def __init__(self, score_thresh=0.05, nms=0.5, detections_per_img=100, box_coder=None, cls_agnostic_bbox_reg=False,
                 bbox_aug_enabled=False):
        super().__init__()
        self.score_thresh, self.nms, self.detections_per_img = score_thresh, nms, detections_per_img

    def prepare_boxlist(self, boxes, scores, image_shape):
        bl = BoxList(boxes.reshape(-1, 4), image_shape, mode="xyxy")
        bl.add_field("scores", scores.reshape(-1))
        return bl


def make_inputs(seed=23, R=150, num_classes=5):
    g = torch.Generator().manual_seed(seed)
    xy = torch.stack((torch.randint(0, 560, (R,), generator=g), torch.randint(0, 400, (R,), generator=g)), 1).float()
    wh = torch.randint(15, 160, (R, 2), generator=g).float()
    proposals = torch.cat((xy, xy + wh), 1)
    deltas = torch.randn(R, 4 * num_classes, generator=g) * 0.15
    scores = torch.rand(R, num_classes, generator=g) * 3.2 - 2.6               # about a fifth above the -2 threshold... of each class
    scores[:, 0] = -1.0
    scores = (scores * 4096).round() / 4096                                    # exactly representable, tie-free enough
    return proposals, deltas, scores


def main():
    if not os.path.isdir(os.path.join(base.REF, "src")):
        raise SystemExit("reference tree not found at %s" % base.REF)
    names = ["maskrcnn_benchmark", "maskrcnn_benchmark.structures", "maskrcnn_benchmark.structures.bounding_box",
             "maskrcnn_benchmark.structures.boxlist_ops", "mrcnn_modified", "mrcnn_modified.modeling",
             "mrcnn_modified.modeling.roi_heads", "mrcnn_modified.modeling.roi_heads.box_head",
             "mrcnn_modified.modeling.roi_heads.box_head.inference"]
    mods = {n: types.ModuleType(n) for n in names}
    mods["maskrcnn_benchmark.structures.bounding_box"].BoxList = BoxList
    mods["maskrcnn_benchmark.structures.boxlist_ops"].boxlist_nms = boxlist_nms
    mods["maskrcnn_benchmark.structures.boxlist_ops"].cat_boxlist = cat_boxlist
    mods["mrcnn_modified.modeling.roi_heads.box_head.inference"].PostProcessor = PostProcessor
    sys.modules.update(mods)
    sys.path.insert(0, os.path.join(base.REF, "src"))
    arrays = {}
    with base.cuda_is_cpu(), contextlib.redirect_stdout(io.StringIO()):
        path = os.path.join(base.REF, "src", "modules", "accuracy-evaluator", "OnlineDetectionPostProcessor.py")
        spec = importlib.util.spec_from_file_location("ref_post", path)
        post = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(post)
        proposals, deltas, scores = make_inputs()
        arrays.update(in_proposals=proposals.numpy(), in_deltas=deltas.numpy(), in_scores=scores.numpy())
        for tag, (thr, nms, k) in {"k40": (-2.0, 0.3, 40), "k100": (-2.0, 0.3, 100), "tight": (-0.2, 0.5, 10),
                                   "none": (5.0, 0.3, 100)}.items():
            pp = post.OnlineDetectionPostProcessor(score_thresh=thr, nms=nms, detections_per_img=k)
            res = pp.forward((scores.clone(), deltas.clone()), [BoxList(proposals.clone(), (640, 480))], 5, (640, 480))
            arrays["%s_boxes" % tag] = res.bbox.numpy()
            arrays["%s_scores" % tag] = res.get_field("scores").numpy()
            arrays["%s_labels" % tag] = res.get_field("labels").numpy()
            arrays["%s_params" % tag] = np.asarray([thr, nms, k], dtype=np.float64)
        # first-party IoU twin (one ground-truth box against all proposals)
        ev_path = os.path.join(base.REF, "src", "modules", "feature-extractor", "mrcnn_modified", "utils", "evaluations.py")
        spec = importlib.util.spec_from_file_location("ref_evaluations", ev_path)
        evm = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(evm)
        gt = torch.tensor([100.0, 80.0, 260.0, 240.0])
        arrays["iou_gt"] = gt.numpy()
        arrays["iou_twin"] = evm.compute_overlap_torch(gt, proposals.clone()).numpy()
    np.savez_compressed(os.path.join(HERE, "reference_post.npz"), **arrays)
    print("wrote reference_post.npz;", {k: v.shape for k, v in arrays.items() if k.endswith("_scores")})


if __name__ == "__main__":
    main()
