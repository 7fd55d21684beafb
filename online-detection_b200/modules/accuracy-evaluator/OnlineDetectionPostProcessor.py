"""Drop-in for src/modules/accuracy-evaluator/OnlineDetectionPostProcessor.py (:9-79) and its
`_standalone` twin (:10-103): decode -> clip -> `score > thresh` -> per-class NMS (+1 box convention,
maskrcnn-benchmark boxlist_nms semantics) -> kthvalue top-K, as ONE libodf call per image
(odf_decode_boxes + odf_detect_postprocess, include/odf.h) instead of a Python loop over classes with a
`.cpu()` kthvalue.  The reference class derives from maskrcnn-benchmark's PostProcessor only to inherit
its constructor fields; the same keyword arguments are accepted here."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import _paths  # noqa: E402,F401
from odf import ops  # noqa: E402

try:
    from maskrcnn_benchmark.structures.bounding_box import BoxList
except Exception:  # noqa: BLE001
    from boxlist import BoxList


class OnlineDetectionPostProcessor(torch.nn.Module):
    def __init__(self, score_thresh=0.05, nms=0.5, detections_per_img=100, box_coder=None,
                 cls_agnostic_bbox_reg=False, bbox_aug_enabled=False):
        super().__init__()
        self.score_thresh = score_thresh
        self.nms = nms
        self.detections_per_img = detections_per_img
        self.box_coder = box_coder
        self.cls_agnostic_bbox_reg = cls_agnostic_bbox_reg
        self.bbox_aug_enabled = bbox_aug_enabled

    # reference :12-33 — x = (class scores [R x C], box deltas [R x 4C]); proposals = [BoxList]
    def forward(self, x, proposals=None, num_classes=None, img_size=None):
        if proposals is None or isinstance(proposals, int):
            return self.forward_standalone(x, proposals if num_classes is None else num_classes)
        cls_scores, bbox_pred = x
        props = proposals[0]
        if img_size is not None and hasattr(props, "resize") and tuple(props.size) != tuple(img_size):
            props = props.resize(img_size)
        size = props.size if img_size is None else img_size
        dev = "cuda"
        ex = props.bbox.to(dev, torch.float32)
        refined = ops.decode_boxes(ex, bbox_pred.to(dev, torch.float32), size[0], size[1])
        boxlist = self.prepare_boxlist(refined, cls_scores.to(dev, torch.float32), size)
        boxlist = boxlist.clip_to_image(remove_empty=False)
        return self.filter_results(boxlist, num_classes)

    # `_standalone` flavour (:11-60): boxes = [BoxList with field "scores"], already decoded
    def forward_standalone(self, boxes, num_classes):
        results = []
        for box in boxes:
            bb = box.bbox
            if self.cls_agnostic_bbox_reg:
                bb = bb.repeat(1, num_classes)
            boxlist = self.prepare_boxlist(bb.to("cuda", torch.float32), box.get_field("scores").to("cuda", torch.float32), box.size)
            boxlist = boxlist.clip_to_image(remove_empty=False)
            results.append(self.filter_results(boxlist, num_classes))
        return results

    def prepare_boxlist(self, boxes, scores, image_shape):
        boxes = boxes.reshape(-1, 4)
        scores = scores.reshape(-1)
        boxlist = BoxList(boxes, image_shape, mode="xyxy")
        boxlist.add_field("scores", scores)
        return boxlist

    # reference :35-79
    def filter_results(self, boxlist, num_classes):
        boxes = boxlist.bbox.reshape(-1, num_classes * 4)
        scores = boxlist.get_field("scores").reshape(-1, num_classes)
        ob, os_, ol, orr = ops.detect_postprocess(boxes, scores, self.score_thresh, self.nms, self.detections_per_img)
        result = BoxList(ob, boxlist.size, mode="xyxy")
        result.add_field("scores", os_)
        result.add_field("labels", ol)
        result.add_field("rois", orr)          # source RoI of every detection (extra, not in the reference)
        return result
