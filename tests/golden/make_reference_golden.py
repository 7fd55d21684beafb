"""Golden vectors produced by RUNNING THE REFERENCE'S OWN FIRST-PARTY CODE in this container (CPU).

The reference's on-line learners (`/root/reference/src`) are plain Python; what keeps them from importing here is
three absent third-party packages and a hard-coded "cuda" device.  This script imports the UNMODIFIED reference files

    src/py_od_utils.py
    src/modules/region-classifier/FALKONWrapper_with_centers_selection_incore.py   (+ MyCenterSelector.py)
    src/modules/region-classifier/OnlineRegionClassifier_incore.py
    src/modules/region-refiner/region_refiner.py  (+ region_refiner_trainer/, region_predictor/)
    src/modules/feature-extractor/mrcnn_modified/modeling/roi_heads/box_head/roi_box_predictors.py   (inference head)

behind three shims that are defined below and nowhere else:

  * `falkon` (un-vendored, pinned at 0d96c685, not installable offline): a stub package whose `InCoreFalkon` /
    `Falkon` / `kernels.GaussianKernel` / `options.FalkonOptions` record how the reference calls them and compute
    with the CPU oracle (`oracle/falkon_oracle.py`, fp64 arithmetic with falkon's fp32 epsilons).  The ARITHMETIC of the
    fit therefore stays "parity unpinned" (it is the oracle's); everything AROUND it — centre selection with the
    reference's RNG calls, the minibootstrap hard/easy negative selection, z-scoring, label construction, the test-time
    score layout, the per-class model list — is the reference's own control flow executing here.
  * `maskrcnn_benchmark.structures.bounding_box.BoxList`: a 20-line container with the fields the files touch.
  * `device='cuda'` / `.to('cuda')` are mapped to the CPU, and `torch.eig` (removed from torch >= 2) is restated with
    `torch.linalg.eig` in the removed function's return format.

The outputs are committed as `tests/golden/reference_flow.npz` (+ `.json`); `tests/test_reference_golden.py` checks the
oracle's restatements against them on the CPU and the product modules against them on the GPU.

    python tests/golden/make_reference_golden.py          # needs /root/reference (not present on the GPU box)
"""
import contextlib
import io
import json
import os
import sys
import types

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("ODF_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import falkon_oracle as orc  # noqa: E402

CALLS = []          # how the reference called the third-party package


# ------------------------------------------------------------------------------------------ shims
def install_falkon_stub():
    falkon = types.ModuleType("falkon")
    kernels = types.ModuleType("falkon.kernels")
    options = types.ModuleType("falkon.options")

    class GaussianKernel:
        def __init__(self, sigma):
            self.sigma = float(sigma)

        def mmv(self, X1, X2, v, out=None):
            res = orc.mmv(X1, X2, v, self.sigma).to(torch.float32)
            if out is not None:
                out.copy_(res)
                return out
            return res

    class FalkonOptions:
        def __init__(self, **kw):
            self.kw = dict(kw)

    class _Falkon:
        flavour = "Falkon"

        def __init__(self, kernel, penalty, M, center_selection=None, maxiter=20, options=None, **kw):
            self.kernel, self.penalty, self.M, self.maxiter = kernel, penalty, M, maxiter
            self.center_selection, self.options = center_selection, options
            CALLS.append({"ctor": self.flavour, "penalty": float(penalty), "M": int(M), "maxiter": int(maxiter),
                          "sigma": kernel.sigma, "options": dict(options.kw) if options is not None else None,
                          "extra": sorted(kw)})

        def fit(self, X, Y):
            Y2 = Y.reshape(Y.shape[0], -1)
            C = self.center_selection.select(X, None)
            CALLS.append({"fit": [int(X.shape[0]), int(X.shape[1])], "y_shape": list(Y.shape), "centres": int(C.shape[0])})
            alpha = orc.falkon_fit(X, Y2, C, self.kernel.sigma, self.penalty, maxiter=self.maxiter, dtype=torch.float64,
                                   eps_pc=1e-5, eps_cg=1e-7)
            self.ny_points_ = C.clone()
            self.alpha_ = alpha.to(torch.float32)
            return self

        def predict(self, X):
            return orc.falkon_predict(X, self.ny_points_, self.alpha_.double(), self.kernel.sigma).to(torch.float32)

    class InCoreFalkon(_Falkon):
        flavour = "InCoreFalkon"

    kernels.GaussianKernel = GaussianKernel
    options.FalkonOptions = FalkonOptions
    options.__all__ = ["FalkonOptions"]
    falkon.kernels, falkon.options = kernels, options
    falkon.Falkon, falkon.InCoreFalkon, falkon.FalkonOptions = _Falkon, InCoreFalkon, FalkonOptions
    sys.modules.update({"falkon": falkon, "falkon.kernels": kernels, "falkon.options": options})


def install_boxlist_stub():
    class BoxList:
        def __init__(self, bbox, image_size, mode="xyxy"):
            self.bbox, self.size, self.mode, self.extra_fields = torch.as_tensor(bbox), image_size, mode, {}

        def add_field(self, k, v):
            self.extra_fields[k] = v

        def get_field(self, k):
            return self.extra_fields[k]

        def __len__(self):
            return self.bbox.shape[0]

    names = ["maskrcnn_benchmark", "maskrcnn_benchmark.structures", "maskrcnn_benchmark.structures.bounding_box"]
    mods = [types.ModuleType(n) for n in names]
    mods[2].BoxList = BoxList
    mods[0].structures, mods[1].bounding_box = mods[1], mods[2]
    sys.modules.update(dict(zip(names, mods)))
    return BoxList


@contextlib.contextmanager
def cuda_is_cpu():
    """Map the reference's hard-coded device 'cuda' to the CPU and restate the removed torch.eig."""
    def is_cuda(d):
        return (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda")

    saved = {}
    for name in ("ones", "zeros", "full", "empty", "eye", "tensor"):
        fn = getattr(torch, name)
        saved[name] = fn

        def wrap(*a, _fn=fn, **kw):
            if is_cuda(kw.get("device")):
                kw["device"] = "cpu"
            return _fn(*a, **kw)
        setattr(torch, name, wrap)
    to_orig = torch.Tensor.to

    def to(self, *a, **kw):
        a = tuple("cpu" if is_cuda(x) else x for x in a)
        if is_cuda(kw.get("device")):
            kw["device"] = "cpu"
        return to_orig(self, *a, **kw)
    torch.Tensor.to = to

    def eig(S, eigenvectors=False):
        # torch.eig (removed): eigenvalues as (n, 2) [real, imag]; S is symmetric here, so everything is real
        vals, vecs = torch.linalg.eig(S)
        return torch.stack((vals.real, vals.imag), 1), vecs.real
    had_eig = hasattr(torch, "eig")
    old_eig = getattr(torch, "eig", None)
    torch.eig = eig
    try:
        yield
    finally:
        for name, fn in saved.items():
            setattr(torch, name, fn)
        torch.Tensor.to = to_orig
        if had_eig:
            torch.eig = old_eig
        else:
            del torch.eig


# ------------------------------------------------------------------------------------------ inputs
CFG = {"CHOSEN_CLASSES": ["__background__", "a", "b", "c"],
       "ONLINE_REGION_CLASSIFIER": {"CLASSIFIER": {"sigma": 12, "lambda": 0.001, "M": 60, "kernel_type": "gauss"},
                                    "MINIBOOTSTRAP": {"EASY_THRESH": -0.9, "HARD_THRESH": -0.7}},
       "REGION_REFINER": {"opts": {"lambda": 10}}}


def make_inputs(d=24, T=3, n_pos=70, n_batches=4, batch=120, seed=0):
    g = torch.Generator().manual_seed(seed)
    protos = torch.randn(T + 1, d, generator=g) * 1.5 + 0.4
    positives = [protos[t + 1] + 0.7 * torch.randn(n_pos + 5 * t, d, generator=g) for t in range(T)]
    negatives = []
    for t in range(T):
        bs = []
        for _ in range(n_batches):
            # mostly background, some rows of the OTHER classes (hard negatives exist)
            lab = torch.randint(0, T + 1, (batch,), generator=g)
            lab[lab == t + 1] = 0
            bs.append(protos[lab] + 0.7 * torch.randn(batch, d, generator=g))
        negatives.append(bs)
    n_test = 90
    lab = torch.randint(0, T + 1, (n_test,), generator=g)
    test_feat = protos[lab] + 0.7 * torch.randn(n_test, d, generator=g)
    xy = torch.rand(n_test, 2, generator=g) * torch.tensor([500.0, 350.0])
    wh = 10 + torch.rand(n_test, 2, generator=g) * 120
    boxes = torch.cat((xy, xy + wh), 1).round()
    # regressor set: class id, features, 4-d targets
    n_reg = 300
    C = torch.randint(1, T + 1, (n_reg, 1), generator=g).float()
    Xr = protos[C[:, 0].long()] + 0.7 * torch.randn(n_reg, d, generator=g)
    Wt = torch.randn(d, 4, generator=g) * 0.05
    Yr = Xr @ Wt + 0.05 * torch.randn(n_reg, 4, generator=g) + torch.tensor([0.1, -0.2, 0.05, 0.0])
    deltas = torch.randn(n_test, 4 * (T + 1), generator=g) * 0.2
    return dict(positives=positives, negatives=negatives, test_feat=test_feat, test_boxes=boxes, test_labels=lab,
                reg_C=C, reg_X=Xr, reg_Y=Yr, deltas=deltas)


def clone_lists(positives, negatives):
    return [p.clone() for p in positives], [[b.clone() for b in bs] for bs in negatives]


# ------------------------------------------------------------------------------------------ the reference flow
def main():
    if not os.path.isdir(os.path.join(REF, "src")):
        raise SystemExit("reference tree not found at %s (this script only runs where /root/reference exists)" % REF)
    install_falkon_stub()
    BoxList = install_boxlist_stub()
    src = os.path.join(REF, "src")
    for p in (src, os.path.join(src, "modules"), os.path.join(src, "modules", "region-classifier"),
              os.path.join(src, "modules", "region-refiner")):
        sys.path.insert(0, p)
    cfg_path = os.path.join(HERE, "reference_flow_cfg.yaml")
    with open(cfg_path, "w") as f:
        yaml.dump(CFG, f)

    inp = make_inputs()
    out = {}
    log = io.StringIO()
    with cuda_is_cpu(), contextlib.redirect_stdout(log):
        import py_od_utils as UT                                               # reference src/py_od_utils.py
        import FALKONWrapper_with_centers_selection_incore as ref_falkon        # reference wrapper (in-core flavour)
        import OnlineRegionClassifier_incore as ref_ocr
        from region_refiner import RegionRefiner as RefRegionRefiner

        # a1: feature statistics (global RNG: torch.randint), py_od_utils.py:59-95
        pos, neg = clone_lists(inp["positives"], inp["negatives"])
        torch.manual_seed(11)
        stats = UT.computeFeatStatistics_torch(pos, neg, num_samples=400, features_dim=24, cpu_tensor=True)
        out["stats_mean"], out["stats_std"], out["stats_mean_norm"] = stats["mean"], stats["std"], stats["mean_norm"].reshape(1)

        # a3: centre selection rule with the reference's RNG calls, ...incore.py:87-99
        w = ref_falkon.FALKONWrapper(cfg_path)
        y = torch.cat((torch.ones(100), -torch.ones(400)))
        torch.manual_seed(5)
        out["sel_many_pos"] = torch.tensor(w.compute_indices_selection(y))
        y2 = torch.cat((torch.ones(7), -torch.ones(400)))
        torch.manual_seed(6)
        out["sel_few_pos"] = torch.tensor(w.compute_indices_selection(y2))

        # a5/a12/a13: minibootstrap training and test-time scoring, OnlineRegionClassifier_incore.py:84-219
        torch.manual_seed(12)
        clf = ref_falkon.FALKONWrapper(cfg_path)
        rc = ref_ocr.OnlineRegionClassifier(clf, pos, neg, stats, cfg_path=cfg_path)
        models, caches = rc.trainRegionClassifier(opts={"return_caches": True})
        for i, m in enumerate(models):
            out["model%d_alpha" % i], out["model%d_centres" % i] = m.alpha_, m.ny_points_
            out["cache%d_neg" % i] = caches[i]["neg"]
        test = [{"boxes": inp["test_boxes"].numpy(), "feat": inp["test_feat"].numpy(),
                 "gt": np.zeros(len(inp["test_boxes"])), "img_size": (640, 480)}]
        preds = rc.testRegionClassifier(models, test)
        out["test_scores"] = preds[0].get_field("scores")
        out["zscored_pos0"] = pos[0]                                            # trainRegionClassifier z-scores in place

        # the out-of-core ("--CPU") flavour: FALKONWrapper_with_centers_selection.py + OnlineRegionClassifier.py
        # (features parked on the host, falkon.Falkon with three options, easy-negative pruning also after the last batch)
        import FALKONWrapper_with_centers_selection as ref_falkon_ooc
        import OnlineRegionClassifier as ref_ocr_ooc
        pos2, neg2_ = clone_lists(inp["positives"], inp["negatives"])
        n_calls = len(CALLS)
        torch.manual_seed(12)
        clf2 = ref_falkon_ooc.FALKONWrapper(cfg_path)
        rc2 = ref_ocr_ooc.OnlineRegionClassifier(clf2, pos2, neg2_, stats, cfg_path=cfg_path)
        models2 = rc2.trainRegionClassifier()
        for i, m in enumerate(models2):
            out["ooc_model%d_alpha" % i], out["ooc_model%d_centres" % i] = m.alpha_, m.ny_points_
        ooc_calls = CALLS[n_calls:]
        del CALLS[n_calls:]

        # a14: RLS refiners on the normalised COXY, train_region_refiner.py:25-119 (+ normalize_COXY, py_od_utils.py:105-111)
        COXY = {"C": inp["reg_C"].clone(), "O": None, "X": inp["reg_X"].clone(), "Y": inp["reg_Y"].clone()}
        COXY = UT.normalize_COXY(COXY, stats, cpu=True)
        out["coxy_X_norm"] = COXY["X"]
        rr = RefRegionRefiner(cfg_path)
        reg = rr.trainRegionRefiner(COXY)
        for i, m in enumerate(reg):
            out["rls%d_mu" % i], out["rls%d_T" % i], out["rls%d_Tinv" % i] = m["mu"], m["T"], m["T_inv"]
            out["rls%d_W" % i] = torch.stack([m["Beta"][str(k)]["weights"] for k in range(4)], 1)
            out["rls%d_losses" % i] = torch.stack([m["Beta"][str(k)]["losses"] for k in range(4)], 1)

        # f-2: region predictor (apply regressors per class, decode with np.spacing(1), clip), predict_regions.py:16-80
        bl_pred = BoxList(inp["test_boxes"].clone(), (640, 480))
        feats = [{"feat": inp["test_feat"].numpy(), "gt": np.zeros(len(inp["test_boxes"]))}]
        refined = rr.predict([bl_pred], feats)
        out["refined_boxes"] = refined[0].bbox

        # a10/a11 + RLS apply: the inference head that consumes the models, roi_box_predictors.py:32-160, loaded from its
        # file with a registry stub (mrcnn_modified/modeling/registry.py only wraps maskrcnn_benchmark's Registry)
        import importlib.util

        class _Reg:
            def register(self, name):
                return lambda cls: cls
        for name in ("mrcnn_modified", "mrcnn_modified.modeling"):
            sys.modules.setdefault(name, types.ModuleType(name))
        regmod = types.ModuleType("mrcnn_modified.modeling.registry")
        regmod.ROI_BOX_PREDICTOR = _Reg()
        sys.modules["mrcnn_modified.modeling.registry"] = regmod
        sys.modules["mrcnn_modified.modeling"].registry = regmod
        head_py = os.path.join(src, "modules", "feature-extractor", "mrcnn_modified", "modeling", "roi_heads", "box_head",
                               "roi_box_predictors.py")
        spec = importlib.util.spec_from_file_location("ref_roi_box_predictors", head_py)
        head_mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(head_mod)
        ns = types.SimpleNamespace
        for parallel in (True, False):
            hcfg = ns(MODEL=ns(ROI_BOX_HEAD=ns(NUM_CLASSES=4), CLS_AGNOSTIC_BBOX_REG=False), INFERENCE=ns(PARALLEL_FALKON=parallel))
            torch.manual_seed(0)
            head = head_mod.FastRCNNPredictor(hcfg, 24)
            head.classifiers = [models[0], None, models[2]] if not parallel else list(models)
            head.regressors = reg
            head.stats = {"mean": stats["mean"], "mean_norm": stats["mean_norm"]}
            head.feat_size = 24
            scores_h, bbox_h = head(inp["test_feat"].clone().reshape(-1, 24, 1, 1))
            tag = "par" if parallel else "seq"
            out["head_scores_" + tag], out["head_bbox_" + tag] = scores_h, bbox_h

        # box decode with the +1 convention, py_od_utils.py:247-274
        bl = BoxList(inp["test_boxes"], (640, 480))
        out["decoded"] = UT.decode_boxes_detector(bl, inp["deltas"].clone())

        # shuffle_negatives / load_positives_from_COXY (global RNG: randperm), py_od_utils.py:226-239,276-294
        _, neg2 = clone_lists(inp["positives"], inp["negatives"])
        torch.manual_seed(21)
        sh = UT.shuffle_negatives(neg2, batch_size=100, num_batches=3)
        out["shuffled_0_0"], out["shuffled_2_2"] = sh[0][0], sh[2][2]
        torch.manual_seed(22)
        lp = UT.load_positives_from_COXY({"C": inp["reg_C"].clone(), "X": inp["reg_X"].clone()}, samples_fraction=0.5)
        out["coxy_pos_1"] = lp[1]

    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}
    for k, v in inp.items():
        if torch.is_tensor(v):
            arrays["in_" + k] = v.numpy()
    for t, p in enumerate(inp["positives"]):
        arrays["in_pos%d" % t] = p.numpy()
    for t, bs in enumerate(inp["negatives"]):
        for j, b in enumerate(bs):
            arrays["in_neg%d_%d" % (t, j)] = b.numpy()
    np.savez_compressed(os.path.join(HERE, "reference_flow.npz"), **arrays)
    with open(os.path.join(HERE, "reference_flow.json"), "w") as f:
        json.dump({"cfg": CFG, "third_party_calls": CALLS, "third_party_calls_out_of_core": ooc_calls,
                   "reference_files": ["src/py_od_utils.py",
                                       "src/modules/region-classifier/FALKONWrapper_with_centers_selection_incore.py",
                                       "src/modules/region-classifier/MyCenterSelector.py",
                                       "src/modules/region-classifier/OnlineRegionClassifier_incore.py",
                                       "src/modules/region-classifier/FALKONWrapper_with_centers_selection.py",
                                       "src/modules/region-classifier/OnlineRegionClassifier.py",
                                       "src/modules/region-refiner/region_refiner.py",
                                       "src/modules/region-refiner/region_refiner_trainer/train_region_refiner.py",
                                       "src/modules/region-refiner/region_predictor/predict_regions.py",
                                       "src/modules/feature-extractor/mrcnn_modified/modeling/roi_heads/box_head/roi_box_predictors.py"],
                   "seeds": {"stats": 11, "sel_many_pos": 5, "sel_few_pos": 6, "minibootstrap": 12, "shuffle": 21,
                             "positives_from_coxy": 22},
                   "torch": torch.__version__}, f, indent=1)
    print("wrote reference_flow.npz (%d arrays, %d third-party calls recorded)" % (len(arrays), len(CALLS)))
    print("cache sizes:", [int(arrays["cache%d_neg" % i].shape[0]) for i in range(3)])


if __name__ == "__main__":
    main()
