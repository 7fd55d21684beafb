"""CPU tests of the host logic above the C ABI: drop-in module surface, centre selection,
pickling, option plumbing, and the row-sharded N>1 fit (world_size 2, gloo) with the TEST-ONLY
CPU operator table from tests/cpu_backend.py injected in place of libodf."""
import copy
import io
import os
import sys

import pytest
import torch
import yaml

from oracle import falkon_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(tmp_path, M=40):
    cfg = {"CHOSEN_CLASSES": ["__background__", "a", "b"],
           "ONLINE_REGION_CLASSIFIER": {"CLASSIFIER": {"sigma": 15, "lambda": 0.001, "M": M, "kernel_type": "gauss"},
                                        "MINIBOOTSTRAP": {"EASY_THRESH": -0.9, "HARD_THRESH": -0.7}},
           "ONLINE_SEGMENTATION": {"CLASSIFIER": {"sigma": 10, "lambda": 1e-6, "M": 20},
                                   "MINIBOOTSTRAP": {"EASY_THRESH": -0.9, "HARD_THRESH": -0.7}},
           "REGION_REFINER": {"opts": {"lambda": 1000}},
           "RPN": {"CHOSEN_CLASSES": ["a%d" % i for i in range(3)],
                   "ONLINE_REGION_CLASSIFIER": {"CLASSIFIER": {"lambda": 1e-5, "M": 30},
                                                "MINIBOOTSTRAP": {"EASY_THRESH": -0.9, "HARD_THRESH": -0.7}},
                   "REGION_REFINER": {"opts": {"lambda": 0.01}}}}
    p = tmp_path / "cfg.yaml"
    p.write_text(yaml.dump(cfg))
    return str(p)


def test_wrapper_reads_config_like_the_reference(tmp_path, capsys):
    import FALKONWrapper_with_centers_selection_incore as falkon
    w = falkon.FALKONWrapper(_cfg(tmp_path))
    assert (w.sigma, w.lam, w.nyst_centers, w.maxiter, w.kernel) == (15, 0.001, 40, 20, None)
    w = falkon.FALKONWrapper(_cfg(tmp_path), is_segmentation=True)
    assert (w.sigma, w.lam, w.nyst_centers) == (10, 1e-6, 20)
    w = falkon.FALKONWrapper(_cfg(tmp_path), is_rpn=True)      # sigma missing -> default 5 + message
    assert w.sigma == 5 and w.lam == 1e-5 and "Sigma not given" in capsys.readouterr().out


def test_wrapper_centre_selection_rule(tmp_path):
    import FALKONWrapper_with_centers_selection as falkon
    w = falkon.FALKONWrapper(_cfg(tmp_path, M=100))
    y = torch.cat((torch.ones(300), -torch.ones(500)))
    idx = w.compute_indices_selection(y)
    assert len(idx) == 100 and sum(i < 300 for i in idx) == 50
    y = torch.cat((torch.ones(10), -torch.ones(500)))
    idx = w.compute_indices_selection(y)
    assert sorted(idx[:10]) == list(range(10)) and len(idx) == 100
    assert w.compute_indices_selection(torch.tensor([1.0])) == 0      # bare int, guarded in train()


def test_center_selector():
    from MyCenterSelector import MyCenterSelector
    X = torch.arange(20.0).reshape(10, 2)
    Y = torch.arange(10.0).reshape(10, 1)
    s = MyCenterSelector([3, 3, 7])
    assert s.select(X, None).tolist() == [[6, 7], [6, 7], [14, 15]]
    Xc, Yc = s.select(X, Y)
    assert Yc.flatten().tolist() == [3, 3, 7]


def test_model_is_picklable_and_truthy():
    from odf import GaussianKernel, InCoreFalkon
    m = InCoreFalkon(kernel=GaussianKernel(5.0), penalty=1e-3, M=4)
    m.ny_points_ = torch.randn(4, 8)
    m.alpha_ = torch.randn(4, 1)
    assert m                                    # `if self.classifiers[i]:` in the reference heads
    m2 = copy.deepcopy(m)
    buf = io.BytesIO()
    torch.save([m, None], buf)
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False)[0]
    for k in (m2, m3):
        assert torch.equal(k.ny_points_, m.ny_points_) and k.kernel.sigma == 5.0 and k.M == 4
    m.alpha_ = m.alpha_ * 2                     # attributes stay assignable (falkon_models_to_cuda)
    assert hasattr(m.kernel, "mmv")


def test_region_classifier_surface(tmp_path):
    import OnlineRegionClassifier_incore as ocr

    class Dummy:
        def __init__(self):
            self.trained = []

        def train(self, X, y, sigma=None, lam=None):
            self.trained.append((len(X), int((y > 0).sum()), sigma, lam))
            return {"w": X[y > 0].mean(0) - X[y < 0].mean(0)}

        def predict(self, model, X):
            return (X @ model["w"])[:, None] * 0 - 0.8     # everything is "medium": kept, not hard

    stats = {"mean": torch.zeros(4), "std": torch.ones(4), "mean_norm": torch.tensor(10.0)}
    pos = [torch.ones(6, 4), torch.empty(0)]
    neg = [[torch.randn(9, 4), torch.randn(9, 4)], [torch.randn(5, 4)]]
    clf = Dummy()
    rc = ocr.OnlineRegionClassifier(clf, pos, neg, stats, cfg_path=_cfg(tmp_path))
    models = rc.trainRegionClassifier(output_dir=str(tmp_path))
    assert len(models) == 2 and models[1] is None and models[0] is not None
    assert clf.trained == [(15, 6, 15, 0.001), (15, 6, 15, 0.001)]     # no hard negatives were added
    assert torch.allclose(pos[0], torch.full((6, 4), 2.0))              # z-scored in place: *20/10
    assert "Detector's Online Classifier training time" in open(tmp_path / "result.txt").read()
    models, caches = rc.trainRegionClassifier(opts={"return_caches": True})
    assert len(caches) == 2 and caches[0]["neg"].shape == (9, 4)


# ------------------------------------------------------------------ N > 1 host logic over gloo
def _worker(rank, world, port, out_dir, dist_pc=None):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "online-detection_b200"))
    sys.path.insert(0, ROOT)
    import cpu_backend
    from odf import Falkon, FalkonOptions, GaussianKernel
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    X, c, Y = orc.make_synthetic(1100, 24, 3, seed=0)
    C = X[orc.shared_centres(c, 64, seed=1)]
    lo, hi = (1100 * rank) // world, (1100 * (rank + 1)) // world
    m = Falkon(GaussianKernel(12.0), 1e-4, 64, process_group=None, _ops=cpu_backend,
               options=FalkonOptions(distributed_precond=dist_pc))
    m.fit(X[lo:hi], Y[lo:hi], centres=C if rank == 0 else torch.zeros_like(C) + C)
    torch.save({"alpha": m.alpha_, "times": m.fit_times_}, os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("dist_pc", [None, True])      # None: replicated build at world 2; True: column blocks + all-gather
@pytest.mark.parametrize("world", [2])
def test_row_sharded_fit_matches_single_rank(tmp_path, world, dist_pc):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path), dist_pc), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r), weights_only=False) for r in range(world)]
    assert torch.equal(outs[0]["alpha"], outs[1]["alpha"])           # replicated CG state stays bitwise equal
    assert outs[0]["times"]["N"] == 1100 and outs[0]["times"]["sweeps"] == 23   # 1 RHS + 20 + 2 restarts
    X, c, Y = orc.make_synthetic(1100, 24, 3, seed=0)
    C = X[orc.shared_centres(c, 64, seed=1)]
    ref = orc.falkon_fit(X, Y, C, 12.0, 1e-4, dtype=torch.float64, eps_pc=1e-5, eps_cg=1e-7)
    p = orc.falkon_predict(X, C, outs[0]["alpha"], 12.0)
    pr = orc.falkon_predict(X, C, ref, 12.0)
    assert float((p - pr).abs().max() / pr.abs().max()) < 1e-3


def test_single_rank_host_logic_matches_oracle_and_freezes_on_convergence():
    import cpu_backend
    from odf import Falkon, FalkonOptions, GaussianKernel
    X, c, Y = orc.make_synthetic(600, 16, 2, seed=2)
    C = X[orc.shared_centres(c, 40, seed=1)]
    m = Falkon(GaussianKernel(10.0), 1e-3, 40, _ops=cpu_backend)
    m.fit(X, Y, centres=C)
    ref = orc.falkon_fit(X, Y, C, 10.0, 1e-3, dtype=torch.float64, eps_pc=1e-5, eps_cg=1e-7)
    assert float((m.alpha_.double() - ref).abs().max() / ref.abs().max()) < 1e-3
    # a huge tolerance converges at iteration 1: exactly one CG step may be applied
    m2 = Falkon(GaussianKernel(10.0), 1e-3, 40, options=FalkonOptions(cg_tolerance=1e3), _ops=cpu_backend)
    m2.fit(X, Y, centres=C)
    ref1 = orc.falkon_fit(X, Y, C, 10.0, 1e-3, dtype=torch.float64, eps_pc=1e-5, eps_cg=1e-7, tol=1e3)
    assert m2.fit_times_["cg_iters"] == 1
    assert float((m2.alpha_.double() - ref1).abs().max() / ref1.abs().max()) < 1e-3
