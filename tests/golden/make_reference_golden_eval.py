"""Golden vectors for the detection evaluator: the REFERENCE's own `eval_detection_icw` / `calc_detection_icw_prec_rec` /
`calc_detection_icw_ap` (src/modules/feature-extractor/mrcnn_modified/data/datasets/evaluation/icubworld/icw_eval.py:
227-402, first-party numpy code) run on synthetic detections.  Shims: the BoxList stub of make_reference_golden.py, empty
stand-ins for cv2 / Masker (imported by the file, unused by these functions), and maskrcnn-benchmark's `boxlist_iou`
restated as SURVEY Appendix B recalls it (areas and intersections with the +1 convention) — that one function is
third-party and un-vendored, so its arithmetic is the oracle's `box_iou_plus1`; everything else is the reference's.

Output: tests/golden/reference_eval.npz, checked against oracle.detection_ap by tests/test_reference_golden.py.

    python tests/golden/make_reference_golden_eval.py      # needs /root/reference
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


def make_cases(seed=17, n_img=6, n_cls=5):
    """Per image: ground truth of a few classes (class 3 never has ground truth, class 4 never has detections in some
    images), detections = jittered copies of the ground truth (several per object: duplicates must count as false
    positives), pure false positives, tie-free scores."""
    rng = np.random.RandomState(seed)
    dets, gts = [], []
    for _ in range(n_img):
        g = rng.randint(1, 5)
        xy = np.stack([rng.randint(0, 500, g), rng.randint(0, 350, g)], 1).astype(np.float32)
        wh = rng.randint(20, 130, size=(g, 2)).astype(np.float32)
        gb = np.concatenate([xy, xy + wh], 1)
        gl = rng.choice([1, 2, 4], size=g)
        db, dl = [], []
        for b, l in zip(gb, gl):
            for _k in range(rng.randint(0, 4)):
                db.append(b + rng.randint(-14, 15, size=4))
                dl.append(l if rng.rand() < 0.8 else rng.choice([1, 2, 3]))
        for _k in range(rng.randint(1, 5)):
            x, y = rng.randint(0, 500), rng.randint(0, 350)
            db.append(np.array([x, y, x + rng.randint(20, 130), y + rng.randint(20, 130)]))
            dl.append(rng.choice([1, 2, 3, 4]))
        db = np.asarray(db, dtype=np.float32).reshape(-1, 4)
        dl = np.asarray(dl, dtype=np.int64)
        ds = rng.permutation(len(dl)).astype(np.float32) / max(len(dl), 1) + 0.013 * len(dets)       # tie-free
        dets.append((db, ds, dl))
        gts.append((gb, gl.astype(np.int64)))
    return dets, gts


def main():
    if not os.path.isdir(os.path.join(base.REF, "src")):
        raise SystemExit("reference tree not found at %s" % base.REF)
    BoxList = base.install_boxlist_stub()
    ops = types.ModuleType("maskrcnn_benchmark.structures.boxlist_ops")
    ops.boxlist_iou = lambda a, b: torch.from_numpy(orc.box_iou_plus1(np.asarray(a.bbox), np.asarray(b.bbox)))
    sys.modules["maskrcnn_benchmark.structures.boxlist_ops"] = ops
    sys.modules["maskrcnn_benchmark.structures"].boxlist_ops = ops
    for name in ("cv2", "mrcnn_modified", "mrcnn_modified.modeling", "mrcnn_modified.modeling.roi_heads",
                 "mrcnn_modified.modeling.roi_heads.mask_head", "mrcnn_modified.modeling.roi_heads.mask_head.inference"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mrcnn_modified.modeling.roi_heads.mask_head.inference"].Masker = object
    sys.path.insert(0, os.path.join(base.REF, "src"))
    path = os.path.join(base.REF, "src", "modules", "feature-extractor", "mrcnn_modified", "data", "datasets", "evaluation",
                        "icubworld", "icw_eval.py")
    spec = importlib.util.spec_from_file_location("ref_icw_eval", path)
    ev = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(ev)
    dets, gts = make_cases()
    preds, gtl = [], []
    for (db, ds, dl), (gb, gl) in zip(dets, gts):
        p = BoxList(torch.from_numpy(db), (640, 480))
        p.add_field("labels", torch.from_numpy(dl))
        p.add_field("scores", torch.from_numpy(ds))
        g = BoxList(torch.from_numpy(gb), (640, 480))
        g.add_field("labels", torch.from_numpy(gl))
        g.add_field("difficult", torch.zeros(len(gl), dtype=torch.uint8))
        preds.append(p)
        gtl.append(g)
    arrays = {}
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for thr in (0.5, 0.7):
            for m07 in (True, False):
                res = ev.eval_detection_icw(preds, gtl, iou_thresh=thr, use_07_metric=m07)
                tag = "iou%02d_%s" % (int(thr * 10), "voc07" if m07 else "area")
                arrays["ap_" + tag] = np.asarray(res["ap"], dtype=np.float64)
                arrays["map_" + tag] = np.asarray([res["map"]], dtype=np.float64)
        # an image list without any ground truth for the predicted classes / without predictions
        res = ev.eval_detection_icw(preds[:1], gtl[:1], iou_thresh=0.5, use_07_metric=True)
        arrays["ap_single"] = np.asarray(res["ap"], dtype=np.float64)
    for i, ((db, ds, dl), (gb, gl)) in enumerate(zip(dets, gts)):
        arrays["in_det_boxes%d" % i], arrays["in_det_scores%d" % i], arrays["in_det_labels%d" % i] = db, ds, dl
        arrays["in_gt_boxes%d" % i], arrays["in_gt_labels%d" % i] = gb, gl
    arrays["n_img"] = np.asarray([len(dets)])
    np.savez_compressed(os.path.join(HERE, "reference_eval.npz"), **arrays)
    print("wrote reference_eval.npz;", {k: np.round(v, 4).tolist() for k, v in arrays.items() if k.startswith(("ap_", "map_"))})


if __name__ == "__main__":
    main()
