"""Sharded loading of the reference's feature caches (odf/shards.py): every file goes to exactly one rank, the union of the
shards is the reference loader's row set, and a row-sharded fit over the shards (gloo, world 2, CPU operator table) gives
the single-rank fit's alpha."""
import os
import socket
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "online-detection_b200"), os.path.join(ROOT, "online-detection_b200", "modules"),
          os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
import format_fixture  # noqa: E402
from oracle import falkon_oracle as orc  # noqa: E402


def _rows_sorted(t):
    if t.numel() == 0:
        return t.reshape(0, 0)
    key = t[:, 0] * 1e3 + t[:, 1]
    return t[key.argsort()]


def test_every_file_has_one_owner_and_the_union_is_the_reference_loader(tmp_path):
    import py_od_utils as UT
    from odf import shards
    format_fixture.build(str(tmp_path))
    for name, is_segm in (("detector_feats", False), ("RPN_feats", False), ("segm_feats", True)):
        fdir = os.path.join(str(tmp_path), name)
        full_pos, full_neg = UT.load_features_classifier(fdir, is_segm=is_segm)
        for world in (1, 2, 3):
            plan = shards.plan_files(fdir, world)
            assert set(plan.values()) <= set(range(world)) and len(plan) == len(os.listdir(fdir))
            assert plan == shards.plan_files(fdir, world)                                  # deterministic
            parts = [shards.load_classifier_shard(fdir, r, world, is_segm=is_segm) for r in range(world)]
            for c in range(len(full_pos)):
                got = [p[0][c] for p in parts if p[0][c].numel() > 0]
                got = torch.cat(got) if got else torch.empty((0, 0))
                assert torch.equal(_rows_sorted(got), _rows_sorted(full_pos[c]))
                ref_neg = full_neg[c] if is_segm else (torch.cat(full_neg[c]) if len(full_neg[c]) else torch.empty((0,)))
                got = []
                for p in parts:
                    nb = p[1][c]
                    got += [nb] if torch.is_tensor(nb) else list(nb)
                got = [g for g in got if g.numel() > 0]
                got = torch.cat(got) if got else torch.empty((0, 0))
                assert torch.equal(_rows_sorted(got), _rows_sorted(ref_neg))
        # two ranks: no class's negatives all on one rank when it has >= 2 files
        plan = shards.plan_files(fdir, 2)
        for c in range(shards.n_classes(fdir)):
            owners = {r for p, r in plan.items() if os.path.basename(p).startswith("negatives_cl_%d_" % c)}
            n_files = sum(os.path.basename(p).startswith("negatives_cl_%d_" % c) for p in plan)
            assert n_files < 2 or owners == {0, 1}
    # regressor batches
    rdir = os.path.join(str(tmp_path), "reg_feats")
    full = UT.load_features_regressor(rdir)
    parts = [shards.load_regressor_shard(rdir, r, 2) for r in range(2)]
    assert torch.equal(torch.cat([p["X"] for p in parts]), full["X"]) and torch.equal(torch.cat([p["C"] for p in parts]), full["C"])
    empty = shards.load_regressor_shard(rdir, 2, 3)
    assert empty["X"].shape == (0, full["X"].shape[1]) and empty["Y"].shape == (0, 4)


def _write_cache(root, d=16):
    """Two classes, several batches: separable blobs so that a fit means something."""
    g = torch.Generator().manual_seed(3)
    protos = torch.randn(3, d, generator=g) * 2
    os.makedirs(root, exist_ok=True)
    for c in range(2):
        for b in range(2):
            torch.save(protos[c + 1] + 0.5 * torch.randn(40 + 7 * b, d, generator=g), os.path.join(root, "positives_cl_%d_batch_%d" % (c, b)))
        for b in range(4):
            torch.save(protos[0] + 0.5 * torch.randn(90 + 11 * b + c, d, generator=g), os.path.join(root, "negatives_cl_%d_batch_%d" % (c, b)))


def _worker(rank, world, port, root, out_dir):
    import torch.distributed as dist
    import cpu_backend
    from odf import Falkon, GaussianKernel, shards
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    pos, neg = shards.load_classifier_shard(root, rank, world)
    X, y = shards.class_fit_rows(pos, neg, 0)
    n_glob = shards.global_count(X.shape[0], dist)
    C = torch.load(os.path.join(out_dir, "centres.pt"))
    m = Falkon(GaussianKernel(6.0), 1e-3, C.shape[0], process_group=None, _ops=cpu_backend)
    m.fit(X, y, centres=C)
    torch.save({"alpha": m.alpha_, "n_local": X.shape[0], "n_glob": n_glob, "N": m.fit_times_["N"]}, os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_row_sharded_fit_over_the_file_shards_matches_the_single_rank_fit(tmp_path):
    import torch.multiprocessing as mp
    import py_od_utils as UT
    import cpu_backend
    from odf import Falkon, GaussianKernel, shards
    root = os.path.join(str(tmp_path), "detector_feats")
    _write_cache(root)
    pos, neg = UT.load_features_classifier(root)
    X, y = shards.class_fit_rows(pos, neg, 0)
    C = X[torch.randperm(X.shape[0], generator=torch.Generator().manual_seed(1))[:48]].contiguous()
    torch.save(C, os.path.join(str(tmp_path), "centres.pt"))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, root, str(tmp_path)), nprocs=2, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r), weights_only=False) for r in range(2)]
    assert torch.equal(outs[0]["alpha"], outs[1]["alpha"])
    assert outs[0]["n_local"] + outs[1]["n_local"] == X.shape[0] == outs[0]["n_glob"] == outs[0]["N"]
    assert min(outs[0]["n_local"], outs[1]["n_local"]) > 0.35 * X.shape[0]                 # balanced by file size
    single = Falkon(GaussianKernel(6.0), 1e-3, 48, _ops=cpu_backend)
    single.fit(X, y, centres=C)
    # same rows in a different order and partition: alpha agrees to rounding
    ref = orc.falkon_predict(X, C, single.alpha_, 6.0)
    got = orc.falkon_predict(X, C, outs[0]["alpha"], 6.0)
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-3
