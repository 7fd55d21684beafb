"""Deterministic on-disk feature caches in the reference's file formats (torch.save pickles named
positives_cl_{c}_batch_{b}, negatives_cl_{c}_batch_{b}, reg_{x,c,y}_batch_{b}; src/py_od_utils.py:120-224), shared by
tests/golden/make_reference_golden_formats.py (which runs the REFERENCE's loaders on them) and by
tests/test_reference_golden.py (which runs the product's drop-in loaders on the same files)."""
import os

import torch
import yaml

SEEDS = {"sample_ratio": 41, "shuffle_detector": 42, "shuffle_rpn": 43, "segm_sample": 44, "reg_fraction": 45}
FEAT_CFG = {"MINIBOOTSTRAP": {"DETECTOR": {"SHUFFLE_NEGATIVES": True, "ITERATIONS": 3, "BATCH_SIZE": 40},
                              "RPN": {"SHUFFLE_NEGATIVES": True, "ITERATIONS": 2, "BATCH_SIZE": 55}}}


def build(root):
    """Writes <root>/detector_feats, <root>/RPN_feats, <root>/segm_feats, <root>/reg_feats and <root>/feat_cfg.yaml."""
    g = torch.Generator().manual_seed(7)
    d = 12
    layout = {"detector_feats": {0: (2, 2), 1: (0, 3), 2: (1, 1)},          # class -> (positive batches, negative batches)
              "RPN_feats": {0: (1, 2), 1: (2, 2)},
              "segm_feats": {0: (2, 2), 1: (1, 0), 2: (1, 3)}}
    for name, classes in layout.items():
        p = os.path.join(root, name)
        os.makedirs(p, exist_ok=True)
        for c, (n_pos, n_neg) in classes.items():
            for b in range(n_pos):
                torch.save(torch.randn(15 + 3 * b + c, d, generator=g), os.path.join(p, "positives_cl_%d_batch_%d" % (c, b)))
            for b in range(n_neg):
                torch.save(torch.randn(30 + 5 * b + 2 * c, d, generator=g), os.path.join(p, "negatives_cl_%d_batch_%d" % (c, b)))
    p = os.path.join(root, "reg_feats")
    os.makedirs(p, exist_ok=True)
    for b in range(2):
        n = 25 + 10 * b
        torch.save(torch.randn(n, d, generator=g), os.path.join(p, "reg_x_batch_%d" % b))
        torch.save(torch.randint(1, 4, (n, 1), generator=g).float(), os.path.join(p, "reg_c_batch_%d" % b))
        torch.save(torch.randn(n, 4, generator=g), os.path.join(p, "reg_y_batch_%d" % b))
    cfg = os.path.join(root, "feat_cfg.yaml")
    with open(cfg, "w") as f:
        yaml.dump(FEAT_CFG, f)
    return cfg


def masks():
    g = torch.Generator().manual_seed(8)
    a = (torch.rand(5, 9, 11, generator=g) > 0.6).numpy()
    b = (torch.rand(4, 9, 11, generator=g) > 0.5).numpy()
    a[4] = False
    b[3] = False                                                           # an empty pair: 0 / 0
    return a, b


def flatten(prefix, positives, negatives, out):
    """positives: list of tensors; negatives: list of (list of tensors | tensor) -> npz-friendly dict entries."""
    out[prefix + "_n_classes"] = torch.tensor([len(positives)])
    for i, p in enumerate(positives):
        out["%s_pos%d" % (prefix, i)] = p
    for i, nb in enumerate(negatives):
        if torch.is_tensor(nb):
            out["%s_neg%d" % (prefix, i)] = nb
        else:
            out["%s_neg%d_n" % (prefix, i)] = torch.tensor([len(nb)])
            for j, b in enumerate(nb):
                out["%s_neg%d_%d" % (prefix, i, j)] = b
    return out


def run_all(UT, root, cfg):
    """The call sequence both sides execute: returns a flat dict of tensors / arrays."""
    out = {}
    det, rpn, seg, reg = (os.path.join(root, n) for n in ("detector_feats", "RPN_feats", "segm_feats", "reg_feats"))
    flatten("det", *UT.load_features_classifier(det), out)
    flatten("det_cpu", *UT.load_features_classifier(det, cpu_tensor=True), out)
    torch.manual_seed(SEEDS["sample_ratio"])
    flatten("det_half", *UT.load_features_classifier(det, sample_ratio=0.5), out)
    torch.manual_seed(SEEDS["shuffle_detector"])
    flatten("det_shuffled", *UT.load_features_classifier(det, cfg_feature_extraction=cfg), out)
    torch.manual_seed(SEEDS["shuffle_rpn"])
    flatten("rpn_shuffled", *UT.load_features_classifier(rpn, cfg_feature_extraction=cfg), out)
    flatten("seg", *UT.load_features_classifier(seg, is_segm=True), out)
    torch.manual_seed(SEEDS["segm_sample"])
    flatten("seg_half", *UT.load_features_classifier(seg, is_segm=True, sample_ratio=0.5), out)
    coxy = UT.load_features_regressor(reg)
    out["reg_C"], out["reg_X"], out["reg_Y"] = coxy["C"], coxy["X"], coxy["Y"]
    assert coxy["O"] is None
    torch.manual_seed(SEEDS["reg_fraction"])
    coxy = UT.load_features_regressor(reg, samples_fraction=0.5)
    out["reg_half_C"], out["reg_half_X"], out["reg_half_Y"] = coxy["C"], coxy["X"], coxy["Y"]
    pos, _ = UT.load_features_classifier(det)
    mb = UT.minibatch_positives([pos[0].clone(), pos[2].clone()], 3)
    for i, parts in enumerate(mb):
        out["minibatch%d_n" % i] = torch.tensor([len(parts)])
        for j, part in enumerate(parts):
            out["minibatch%d_%d" % (i, j)] = part
    a, b = masks()
    out["mask_iou"] = UT.mask_iou(a, b)
    out["zscores"] = UT.zScores(pos[0], pos[0].mean(0), pos[0].norm(dim=1).mean())
    out["zscores_30"] = UT.zScores(pos[0], pos[0].mean(0), pos[0].norm(dim=1).mean(), target_norm=30)
    return out
