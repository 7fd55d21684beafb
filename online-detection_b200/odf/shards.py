"""Sharded loading of the reference's on-disk feature caches straight into per-GPU row shards.

The reference writes its extracted features as `torch.save` pickles, one file per (kind, class, batch):
    positives_cl_{c}_batch_{b}   negatives_cl_{c}_batch_{b}          (src/py_od_utils.py:120-199, load_features_classifier)
    reg_x_batch_{b}  reg_c_batch_{b}  reg_y_batch_{b}               (src/py_od_utils.py:202-224, load_features_regressor)
and `load_features_classifier` reads ALL of them into one process.  A row-sharded FALKON fit (odf.Falkon with a process
group: rows split over the GPUs of the box, one all-reduce of the M x T partial per sweep) needs only an arbitrary
PARTITION of the rows, so here every file is read by exactly ONE rank:

    plan_files()              a deterministic file -> rank map, the same on every rank (a pure function of the
                              directory listing and the file sizes): greedy largest-first onto the least loaded rank,
                              per class, so that every class's rows are balanced over the ranks
    load_classifier_shard()   (positives, negatives) in the reference loader's structure, holding this rank's files
                              only (classes without a local file get an empty (0 x d) tensor / an empty list)
    load_regressor_shard()    the COXY dictionary of this rank's reg_* batches
    class_fit_rows()          X (n_local x d), y (n_local,) = +-1 of one class for Falkon.fit(..., process_group=...)

Nothing here touches the GPU unless `device` is given (the tensors are then moved as they are read: one file in host
memory at a time).  The union of the shards is the reference loader's row set (tests/test_shards.py)."""
import glob
import os
import re

import torch

_CLS_RE = re.compile(r"^(positives|negatives)_cl_(\d+)_batch_(\d+)$")
_REG_RE = re.compile(r"^reg_x_batch_(\d+)$")


def _listing(features_dir):
    """[(kind, class, batch, path, bytes)] of the classifier cache files, sorted."""
    out = []
    for p in glob.glob(os.path.join(features_dir, "*_cl_*_batch_*")):
        m = _CLS_RE.match(os.path.basename(p))
        if m:
            out.append((m.group(1), int(m.group(2)), int(m.group(3)), p, os.path.getsize(p)))
    return sorted(out)


def plan_files(features_dir, world):
    """{path: rank}.  Per (kind, class): files in decreasing size, each to the rank with the fewest bytes of that kind and
    class so far (ties -> the rank with the fewest bytes overall, then the lowest rank).  Deterministic."""
    plan, total = {}, [0] * world
    groups = {}
    for kind, cls, batch, path, size in _listing(features_dir):
        groups.setdefault((kind, cls), []).append((size, batch, path))
    for key in sorted(groups):
        load = [0] * world
        for size, batch, path in sorted(groups[key], key=lambda t: (-t[0], t[1])):
            r = min(range(world), key=lambda k: (load[k], total[k], k))
            plan[path] = r
            load[r] += size
            total[r] += size
    return plan


def n_classes(features_dir):
    lst = _listing(features_dir)
    return 1 + max(c for _k, c, _b, _p, _s in lst) if lst else 0


def load_classifier_shard(features_dir, rank, world, is_segm=False, device=None):
    """This rank's part of `load_features_classifier(features_dir, is_segm)`: positives = list[T] of (P_i x d) tensors,
    negatives = list[T] of lists of batch tensors (or of single tensors when is_segm), from the files plan_files() gives
    this rank.  Also returns the global row counts {'positives': [T], 'negatives': [T]} when they can be had without
    reading the other ranks' files -- they cannot (pickles carry no header), so counts are local; all-reduce them."""
    plan = plan_files(features_dir, world)
    T = n_classes(features_dir)
    pos = [[] for _ in range(T)]
    neg = [[] for _ in range(T)]
    for kind, cls, batch, path, _size in _listing(features_dir):
        if plan[path] != rank:
            continue
        t = torch.load(path)
        if device is not None:
            t = t.to(device)
        (pos if kind == "positives" else neg)[cls].append(t)

    def cat(parts):
        return torch.cat(parts) if parts else torch.empty((0,))
    positives = [cat(p) for p in pos]
    negatives = [cat(n) for n in neg] if is_segm else neg
    return positives, negatives


def load_regressor_shard(features_dir, rank, world, device=None):
    """This rank's part of `load_features_regressor`: batches b with b % world == rank (the three files of a batch go
    together).  Ranks without a batch get empty tensors of the right width (d is read from batch 0's x header row)."""
    ids = sorted(int(_REG_RE.match(os.path.basename(p)).group(1)) for p in glob.glob(os.path.join(features_dir, "reg_x_batch_*"))
                 if _REG_RE.match(os.path.basename(p)))
    X, C, Y = [], [], []
    for b in ids:
        if b % world != rank:
            continue
        parts = [torch.load(os.path.join(features_dir, "reg_%s_batch_%d" % (k, b))) for k in ("x", "c", "y")]
        if device is not None:
            parts = [t.to(device) for t in parts]
        X.append(parts[0])
        C.append(parts[1])
        Y.append(parts[2])
    if not X:
        d = torch.load(os.path.join(features_dir, "reg_x_batch_%d" % ids[0])).shape[1] if ids else 0
        return {"C": torch.empty((0, 1)), "O": None, "X": torch.empty((0, d)), "Y": torch.empty((0, 4))}
    return {"C": torch.cat(C), "O": None, "X": torch.cat(X), "Y": torch.cat(Y)}


def class_fit_rows(positives, negatives, cls, is_segm=False):
    """Local rows and +-1 labels of one class, in the order the reference concatenates them
    (OnlineRegionClassifier.py:100-107: positives first, then the negative batches)."""
    p = positives[cls]
    nb = negatives[cls]
    n = nb if (is_segm or torch.is_tensor(nb)) else (torch.cat(nb) if len(nb) else torch.empty((0,)))
    parts = [t for t in (p, n) if t.numel() > 0]
    if not parts:
        return torch.empty((0, 0)), torch.empty((0,))
    X = torch.cat(parts)
    y = torch.cat((torch.ones(p.shape[0] if p.numel() else 0), -torch.ones(n.shape[0] if n.numel() else 0))).to(X.device)
    return X, y


def global_count(local_count, dist=None, group=None, device=None):
    """Sum of a per-rank row count over the process group (the `N` of a sharded fit is taken care of by Falkon.fit
    itself; this is for callers that need class sizes, e.g. to decide `M` or to skip empty classes consistently)."""
    if dist is None:
        return int(local_count)
    t = torch.tensor([float(local_count)], dtype=torch.float64, device=device)
    dist.all_reduce(t, group=group)
    return int(t.item())
