"""Golden vectors for the OTHER TWO flavours of the reference's on-line learner, produced — like reference_flow.npz — by
RUNNING THE REFERENCE'S OWN first-party code on the CPU behind the shims of make_reference_golden.py (falkon stub =
oracle, BoxList stub, 'cuda' -> CPU):

  * on-line RPN  (BASELINE config 4): `OnlineRegionClassifier(..., is_rpn=True)` — configuration read from the `RPN`
    section, one binary model per anchor class (`num_classes = len(RPN.CHOSEN_CLASSES) + 1`), an anchor without
    positives yields `None`, timing line "RPN's Online Classifier training time" in result.txt
    (src/modules/region-classifier/OnlineRegionClassifier_incore.py:18-53,96-155);
  * on-line segmentation (BASELINE config 3): `is_segmentation=True` — `ONLINE_SEGMENTATION` section, per-pixel
    features, `negatives[i]` a single tensor wrapped in a 1-list (experiments/...oos.py:252-254), so every class is
    fitted exactly once with no hard / easy selection.

Outputs: tests/golden/reference_flow_flavours.npz (+ .json), checked by tests/test_reference_golden.py on the CPU (oracle
restatement of the loop AND the product's drop-in OnlineRegionClassifier module driven by an oracle-backed classifier).

    python tests/golden/make_reference_golden_flavours.py      # needs /root/reference (not present on the GPU box)
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as base  # noqa: E402

CFG = {"CHOSEN_CLASSES": ["__background__", "a", "b", "c"],
       "ONLINE_REGION_CLASSIFIER": {"CLASSIFIER": {"sigma": 12, "lambda": 0.001, "M": 60, "kernel_type": "gauss"},
                                    "MINIBOOTSTRAP": {"EASY_THRESH": -0.9, "HARD_THRESH": -0.7}},
       "ONLINE_SEGMENTATION": {"CLASSIFIER": {"sigma": 9, "lambda": 0.0001, "M": 40, "kernel_type": "gauss"},
                               "MINIBOOTSTRAP": {"EASY_THRESH": -0.85, "HARD_THRESH": -0.6}},
       "REGION_REFINER": {"opts": {"lambda": 10}},
       "RPN": {"CHOSEN_CLASSES": ["anchor0", "anchor1", "anchor2", "anchor3"],
               "ONLINE_REGION_CLASSIFIER": {"CLASSIFIER": {"sigma": 14, "lambda": 0.001, "M": 50, "kernel_type": "gauss"},
                                            "MINIBOOTSTRAP": {"EASY_THRESH": -0.8, "HARD_THRESH": -0.5}},
               "REGION_REFINER": {"opts": {"lambda": 0.01}}}}
SEEDS = {"rpn": 31, "segmentation": 32}


def make_inputs(seed=3):
    g = torch.Generator().manual_seed(seed)
    d_rpn, A = 20, 4                                     # 4 anchor classes, anchor 2 never positive
    protos = torch.randn(A + 1, d_rpn, generator=g) * 1.5 + 0.3
    rpn_pos = [protos[a + 1] + 0.7 * torch.randn(50 + 7 * a, d_rpn, generator=g) if a != 2 else torch.empty((0, d_rpn))
               for a in range(A)]
    rpn_neg = []
    for a in range(A):
        bs = []
        for _ in range(3):
            lab = torch.randint(0, A + 1, (100,), generator=g)
            lab[lab == a + 1] = 0
            bs.append(protos[lab] + 0.7 * torch.randn(100, d_rpn, generator=g))
        rpn_neg.append(bs)
    d_seg, T = 16, 3                                     # per-pixel features, one negative tensor per class in a 1-list
    sp = torch.randn(T + 1, d_seg, generator=g) * 1.2 - 0.2
    seg_pos = [sp[t + 1] + 0.6 * torch.randn(80 + 10 * t, d_seg, generator=g) for t in range(T)]
    seg_neg = []
    for t in range(T):
        lab = torch.randint(0, T + 1, (260,), generator=g)
        lab[lab == t + 1] = 0
        seg_neg.append([sp[lab] + 0.6 * torch.randn(260, d_seg, generator=g)])
    stats = {}
    for name, pos, neg in (("rpn", rpn_pos, rpn_neg), ("seg", seg_pos, seg_neg)):
        allrows = torch.cat([p for p in pos if len(p)] + [b for bs in neg for b in bs], 0)
        stats[name] = {"mean": allrows.mean(0), "std": allrows.std(0), "mean_norm": (allrows - allrows.mean(0)).norm(dim=1).mean()}
    return dict(rpn_pos=rpn_pos, rpn_neg=rpn_neg, seg_pos=seg_pos, seg_neg=seg_neg, stats=stats)


def main():
    if not os.path.isdir(os.path.join(base.REF, "src")):
        raise SystemExit("reference tree not found at %s (this script only runs where /root/reference exists)" % base.REF)
    base.install_falkon_stub()
    base.install_boxlist_stub()
    src = os.path.join(base.REF, "src")
    for p in (src, os.path.join(src, "modules"), os.path.join(src, "modules", "region-classifier")):
        sys.path.insert(0, p)
    inp = make_inputs()
    out, lines = {}, {}
    log = io.StringIO()
    with tempfile.TemporaryDirectory() as tmp, base.cuda_is_cpu(), contextlib.redirect_stdout(log):
        cfg_path = os.path.join(tmp, "cfg.yaml")
        with open(cfg_path, "w") as f:
            yaml.dump(CFG, f)
        import FALKONWrapper_with_centers_selection_incore as ref_falkon
        import OnlineRegionClassifier_incore as ref_ocr
        for tag, kw, pos_key, neg_key, st in (("rpn", {"is_rpn": True}, "rpn_pos", "rpn_neg", "rpn"),
                                             ("seg", {"is_segmentation": True}, "seg_pos", "seg_neg", "seg")):
            pos = [p.clone() for p in inp[pos_key]]
            neg = [[b.clone() for b in bs] for bs in inp[neg_key]]
            n0 = len(base.CALLS)
            torch.manual_seed(SEEDS["rpn" if tag == "rpn" else "segmentation"])
            clf = ref_falkon.FALKONWrapper(cfg_path, **kw)
            rc = ref_ocr.OnlineRegionClassifier(clf, pos, neg, inp["stats"][st], cfg_path=cfg_path, **kw)
            odir = os.path.join(tmp, tag)
            os.makedirs(odir)
            models, caches = rc.trainRegionClassifier(opts={"return_caches": True}, output_dir=odir)
            out[tag + "_n_models"] = torch.tensor([len(models)])
            out[tag + "_is_none"] = torch.tensor([m is None for m in models])
            for i, m in enumerate(models):
                if m is not None:
                    out["%s_model%d_alpha" % (tag, i)], out["%s_model%d_centres" % (tag, i)] = m.alpha_, m.ny_points_
                    out["%s_cache%d_neg" % (tag, i)] = caches[i]["neg"]
            lines[tag] = open(os.path.join(odir, "result.txt")).read()
            lines[tag + "_calls"] = base.CALLS[n0:]
            lines[tag + "_hyper"] = {"sigma": rc.sigma, "lam": rc.lam, "hard": rc.hard_tresh, "easy": rc.easy_tresh,
                                     "M": clf.nyst_centers, "num_classes": rc.num_classes}
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}
    for key in ("rpn_pos", "seg_pos"):
        for i, p in enumerate(inp[key]):
            arrays["in_%s%d" % (key, i)] = p.numpy()
    for key in ("rpn_neg", "seg_neg"):
        for i, bs in enumerate(inp[key]):
            for j, b in enumerate(bs):
                arrays["in_%s%d_%d" % (key, i, j)] = b.numpy()
    for st in ("rpn", "seg"):
        arrays["in_stats_%s_mean" % st] = inp["stats"][st]["mean"].numpy()
        arrays["in_stats_%s_std" % st] = inp["stats"][st]["std"].numpy()
        arrays["in_stats_%s_mean_norm" % st] = inp["stats"][st]["mean_norm"].reshape(1).numpy()
    np.savez_compressed(os.path.join(HERE, "reference_flow_flavours.npz"), **arrays)
    with open(os.path.join(HERE, "reference_flow_flavours.json"), "w") as f:
        json.dump({"cfg": CFG, "seeds": SEEDS, "result_lines": {k: v for k, v in lines.items() if isinstance(v, str)},
                   "hyper": {"rpn": lines["rpn_hyper"], "seg": lines["seg_hyper"]},
                   "third_party_calls": {"rpn": lines["rpn_calls"], "seg": lines["seg_calls"]},
                   "reference_files": ["src/modules/region-classifier/OnlineRegionClassifier_incore.py",
                                       "src/modules/region-classifier/FALKONWrapper_with_centers_selection_incore.py",
                                       "src/modules/region-classifier/MyCenterSelector.py"],
                   "torch": torch.__version__}, f, indent=1)
    print("wrote reference_flow_flavours.npz (%d arrays)" % len(arrays))
    print("rpn models:", arrays["rpn_is_none"].tolist(), " seg models:", arrays["seg_is_none"].tolist())
    print("rpn cache sizes:", [int(arrays["rpn_cache%d_neg" % i].shape[0]) for i in range(4) if "rpn_cache%d_neg" % i in arrays])
    print("result lines:", {k: v.strip() for k, v in lines.items() if isinstance(v, str)})


if __name__ == "__main__":
    main()
