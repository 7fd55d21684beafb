"""Drop-in for src/py_od_utils.py — the helper functions the experiment scripts import
(`computeFeatStatistics_torch, normalize_COXY, falkon_models_to_cuda, load_features_classifier,
load_features_regressor, load_positives_from_COXY, decode_boxes_detector, shuffle_negatives,
mask_iou, zScores, minibatch_positives`), same signatures and return layouts
(reference: src/py_od_utils.py:59-331).  File formats are the reference's: `torch.save` pickles
named positives_cl_{c}_batch_{b}, negatives_cl_{c}_batch_{b}, reg_{x,c,y}_batch_{b}."""
import glob
import math
import os

import numpy as np
import torch
import yaml

_GPU = "cuda"


def computeFeatStatistics_torch(positives, negatives, num_samples=4000, features_dim=2048, cpu_tensor=False,
                                pos_fraction=None):
    """~num_samples rows (pos_fraction of them positives, default 1/10) drawn WITH replacement over
    classes and negative batches -> {'mean' (d), 'std' (d, unbiased), 'mean_norm'} on the GPU
    (reference :59-95; mean_norm is the mean norm of the UN-centred rows)."""
    device = "cpu" if cpu_tensor else _GPU
    print("Computing features statistics")
    pos_fraction = 0.1 if pos_fraction is None else pos_fraction
    neg_fraction = 1 - pos_fraction
    n_cls = len(positives)
    per_class = num_samples / n_cls
    take_pos = math.ceil(per_class * pos_fraction)
    most_batches = max([len(nb) for nb in negatives] + [0])
    take_neg = math.ceil(per_class * neg_fraction / most_batches)
    picked = []
    for i in range(n_cls):
        if len(positives[i]) != 0:
            picked.append(positives[i][torch.randint(len(positives[i]), (take_pos,))].view(-1, features_dim))
        for batch in negatives[i]:
            if len(batch) != 0:
                picked.append(batch[torch.randint(len(batch), (take_neg,))].view(-1, features_dim))
    sample = torch.cat(picked).to(device) if picked else torch.empty((0, features_dim), device=device)
    stats = {"mean": sample.mean(dim=0), "std": sample.std(dim=0), "mean_norm": sample.norm(dim=1).mean()}
    return {k: v.to(_GPU) for k, v in stats.items()}


def computeFeatStatistics(positives, negatives, feature_folder, is_rpn, num_samples=4000):
    """Legacy cached-on-disk variant (reference :8-57): returns (mean, std, mean_norm) CPU tensors,
    loading `<Data>/feat_cache[_RPN]/<folder>/[rpn_]stats` when present, else computing + saving."""
    base = os.path.join(os.path.dirname(__file__), "..", "Data")
    path = os.path.join(base, "feat_cache_RPN" if is_rpn else "feat_cache", feature_folder,
                        "rpn_stats" if is_rpn else "stats")
    try:
        saved = torch.load(path)
        return torch.tensor(saved["mean"]), torch.tensor(saved["std"]), torch.tensor(saved["mean_norm"])
    except Exception:  # noqa: BLE001
        pass
    dim = next(p.shape[-1] for p in positives if len(p) != 0)
    s = computeFeatStatistics_torch([p.cpu() for p in positives], [[b.cpu() for b in nb] for nb in negatives],
                                    num_samples=num_samples, features_dim=dim, cpu_tensor=True) \
        if not torch.cuda.is_available() else computeFeatStatistics_torch(positives, negatives, num_samples, dim)
    mean, std, mean_norm = s["mean"].cpu(), s["std"].cpu(), s["mean_norm"].cpu()
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save({"mean": mean, "std": std, "mean_norm": mean_norm}, path)
    except Exception:  # noqa: BLE001
        pass
    return mean, std, mean_norm


def zScores(feat, mean, mean_norm, target_norm=20):
    feat = torch.as_tensor(feat)
    return (feat - mean) * (target_norm / mean_norm)


def normalize_COXY(COXY, stats, cpu=False):
    mean = stats["mean"].to("cpu") if cpu else stats["mean"]
    COXY["X"] = (COXY["X"] - mean) * (20 / stats["mean_norm"].item())
    return COXY


def falkon_models_to_cuda(models):
    for m in models:
        if m is not None:
            m.ny_points_ = m.ny_points_.to(_GPU)
            m.alpha_ = m.alpha_.to(_GPU)
    return models


def _minibootstrap_shuffle_cfg(cfg_path, features_dir):
    shuffle, batch_size, n_batches = False, 2000, 2
    if cfg_path is None:
        return shuffle, batch_size, n_batches
    with open(cfg_path) as fh:
        params = yaml.load(fh, Loader=yaml.FullLoader)
    mb = params.get("MINIBOOTSTRAP", {})
    for key, tag in (("RPN", "RPN"), ("DETECTOR", "detector")):
        if key in mb and tag in features_dir:
            shuffle = mb[key].get("SHUFFLE_NEGATIVES", shuffle)
            n_batches = mb[key].get("ITERATIONS", n_batches)
            batch_size = mb[key].get("BATCH_SIZE", batch_size)
    return shuffle, batch_size, n_batches


def _load_batches(features_dir, stem, cls):
    n = len(glob.glob(os.path.join(features_dir, "%s_cl_%d_*" % (stem, cls))))
    return [torch.load(os.path.join(features_dir, "%s_cl_%d_batch_%d" % (stem, cls, b))) for b in range(n)]


def _cat_or_empty(parts, cpu_tensor, sample_ratio):
    try:
        if cpu_tensor:
            return torch.cat(parts).to("cpu")
        allrows = torch.cat(parts)
        if sample_ratio < 1:
            allrows = allrows[torch.randint(len(allrows), (int(len(allrows) * sample_ratio),))]
        return allrows
    except Exception:  # noqa: BLE001  (no files for this class)
        return torch.empty((0))


def load_features_classifier(features_dir, is_segm=False, cpu_tensor=False, sample_ratio=1,
                             cfg_feature_extraction=None):
    """-> (positives: list[T] of (P_i x d), negatives: list[T] of list[batches] — or of single
    tensors when is_segm) from the on-disk feature cache (reference :120-200)."""
    left_pos = len(glob.glob(os.path.join(features_dir, "positives_*")))
    left_neg = len(glob.glob(os.path.join(features_dir, "negatives_*")))
    shuffle, bs, nb = _minibootstrap_shuffle_cfg(cfg_feature_extraction, features_dir)
    positives, negatives = [], []
    cls = 0
    while left_pos > 0 or left_neg > 0:
        p = _load_batches(features_dir, "positives", cls)
        left_pos -= len(p)
        positives.append(_cat_or_empty(p, cpu_tensor, sample_ratio))
        n = _load_batches(features_dir, "negatives", cls)
        left_neg -= len(n)
        negatives.append(_cat_or_empty(n, cpu_tensor, sample_ratio) if is_segm else n)
        cls += 1
    if not is_segm and shuffle:
        negatives = shuffle_negatives(negatives, batch_size=bs, num_batches=nb)
    return positives, negatives


def load_features_regressor(features_dir, samples_fraction=1.0):
    """-> COXY = {'C': (n,1) class ids, 'O': None, 'X': (n x d), 'Y': (n x 4)} (reference :202-224)."""
    n_batches = len(glob.glob(os.path.join(features_dir, "reg_x_*")))
    X, C, Y = [], [], []
    for b in range(n_batches):
        c = torch.load(os.path.join(features_dir, "reg_c_batch_%d" % b))
        x = torch.load(os.path.join(features_dir, "reg_x_batch_%d" % b))
        y = torch.load(os.path.join(features_dir, "reg_y_batch_%d" % b))
        if samples_fraction < 1.0:
            keep = torch.randperm(len(c))[:int(len(c) * samples_fraction)]
            c, x, y = c[keep], x[keep], y[keep]
        X.append(x)
        C.append(c)
        Y.append(y)
    return {"C": torch.cat(C), "O": None, "X": torch.cat(X), "Y": torch.cat(Y)}


def load_positives_from_COXY(COXY, del_COXY=False, samples_fraction=1.0):
    positives = []
    for i in range(len(torch.unique(COXY["C"]))):
        ids = torch.where(COXY["C"] == i + 1)[0]
        if samples_fraction < 1.0:
            ids = ids[torch.randperm(len(ids))[:int(len(ids) * samples_fraction)]]
        positives.append(COXY["X"][ids])
        if del_COXY:
            rest = torch.where(COXY["C"] != i + 1)[0]
            COXY["X"] = COXY["X"][rest]
            COXY["C"] = COXY["C"][rest]
    return positives


def minibatch_positives(positives, num_batches):
    for i in range(len(positives)):
        positives[i] = list(torch.split(positives[i], int(len(positives[i]) / num_batches)))
    return positives


def decode_boxes_detector(boxes, bbox_pred):
    """Deltas -> boxes in the legacy +1 convention, clipped to the image (reference :247-274)."""
    ex = boxes.bbox
    w = (ex[:, 2] - ex[:, 0] + 1)[:, None]
    h = (ex[:, 3] - ex[:, 1] + 1)[:, None]
    cx = ex[:, 0:1] + 0.5 * w
    cy = ex[:, 1:2] + 0.5 * h
    pcx = bbox_pred[:, 0::4] * w + cx
    pcy = bbox_pred[:, 1::4] * h + cy
    pw = torch.exp(bbox_pred[:, 2::4]) * w
    ph = torch.exp(bbox_pred[:, 3::4]) * h
    out = torch.zeros_like(bbox_pred)
    out[:, 0::4] = (pcx - 0.5 * pw).clamp(min=0)
    out[:, 1::4] = (pcy - 0.5 * ph).clamp(min=0)
    out[:, 2::4] = (pcx + 0.5 * pw - 1).clamp(max=boxes.size[0] - 1)
    out[:, 3::4] = (pcy + 0.5 * ph - 1).clamp(max=boxes.size[1] - 1)
    return out


def shuffle_negatives(negatives, batch_size=None, num_batches=None):
    out = []
    for per_class in negatives:
        bs = len(per_class[0]) if batch_size is None else batch_size
        pool = torch.cat(per_class)
        nb = math.ceil(len(pool) / bs) if num_batches is None else num_batches
        order = torch.randperm(len(pool))
        out.append([pool[order[min(j * bs, len(order)):min((j + 1) * bs, len(order))]] for j in range(nb)])
    return out


def mask_iou(mask_a, mask_b):
    """IoU between two stacks of boolean masks (N,H,W) x (K,H,W) -> (N,K) float32 (reference
    :297-331), vectorised."""
    if mask_a.shape[1:] != mask_b.shape[1:]:
        raise IndexError
    a = np.asarray(mask_a).reshape(len(mask_a), -1).astype(bool)
    b = np.asarray(mask_b).reshape(len(mask_b), -1).astype(bool)
    inter = (a[:, None, :] & b[None, :, :]).sum(-1)
    union = (a[:, None, :] | b[None, :, :]).sum(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / union).astype(np.float32)
