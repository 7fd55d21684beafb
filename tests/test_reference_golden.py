"""Parity against golden vectors produced by RUNNING THE REFERENCE'S OWN first-party code
(tests/golden/make_reference_golden.py: unmodified py_od_utils.py, FALKONWrapper_with_centers_selection_incore.py,
MyCenterSelector.py, OnlineRegionClassifier_incore.py, region_refiner.py + trainer, executed on the CPU behind a
`falkon` stub that computes with the oracle, a BoxList stub and a cuda->cpu device shim).

CPU tests: the oracle's restatements and the product's torch-only helpers reproduce the reference's outputs (bit-exact
where the work is indices / RNG draws / elementwise fp32, to rounding where it is fp64 linear algebra).
GPU tests: the product modules (drop-in file names, CUDA path through the C ABI) reproduce them too.
The third-party FALKON arithmetic itself is the oracle's in those fixtures — that part stays "parity unpinned"."""
import json
import os

import numpy as np
import pytest
import torch
import yaml

from oracle import falkon_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "reference_flow.npz"))
META = json.load(open(os.path.join(HERE, "golden", "reference_flow.json")))
CFG = META["cfg"]
SEEDS = META["seeds"]
T_CLS, N_BATCH = 3, 4


def t(name):
    return torch.from_numpy(G[name])


def inputs():
    positives = [t("in_pos%d" % i).clone() for i in range(T_CLS)]
    negatives = [[t("in_neg%d_%d" % (i, j)).clone() for j in range(N_BATCH)] for i in range(T_CLS)]
    return positives, negatives


def cfg_file(tmp_path):
    p = tmp_path / "cfg.yaml"
    p.write_text(yaml.dump(CFG))
    return str(p)


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ------------------------------------------------------------------------------------------ CPU: oracle / host helpers
def test_reference_called_falkon_the_way_the_product_assumes():
    """What the reference handed to the third-party package (recorded by the stub): in-core flavour, maxiter 20, the
    six FalkonOptions, y as a flat (N,) vector, M = number of selected centres (<= the configured budget)."""
    calls = META["third_party_calls"]
    ctors = [c for c in calls if "ctor" in c]
    fits = [c for c in calls if "fit" in c]
    assert len(ctors) == len(fits) == T_CLS * N_BATCH
    for c in ctors:
        assert c["ctor"] == "InCoreFalkon" and c["maxiter"] == 20 and c["sigma"] == 12.0 and c["penalty"] == 0.001
        assert c["M"] <= 60 and c["extra"] == []
        assert c["options"] == {"min_cuda_iter_size_32": 0, "min_cuda_iter_size_64": 0, "keops_active": "no",
                                "min_cuda_pc_size_32": 0, "min_cuda_pc_size_64": 0, "store_kernel_d_threshold": 250}
    for c, f in zip(ctors, fits):
        assert f["y_shape"] == [f["fit"][0]] and f["centres"] == c["M"]


def test_reference_out_of_core_flavour_calls():
    """The `--CPU` flavour (FALKONWrapper_with_centers_selection.py + OnlineRegionClassifier.py): falkon.Falkon, three
    options, no explicit maxiter (falkon's default 20), and — same RNG stream, same arithmetic — the same models."""
    calls = META["third_party_calls_out_of_core"]
    ctors = [c for c in calls if "ctor" in c]
    assert len(ctors) == T_CLS * N_BATCH
    for c in ctors:
        assert c["ctor"] == "Falkon" and c["maxiter"] == 20
        assert c["options"] == {"min_cuda_iter_size_32": 0, "min_cuda_iter_size_64": 0, "keops_active": "no"}
    for i in range(T_CLS):
        assert np.array_equal(G["ooc_model%d_alpha" % i], G["model%d_alpha" % i])
        assert np.array_equal(G["ooc_model%d_centres" % i], G["model%d_centres" % i])


def test_oracle_centre_selection_matches_reference_rng():
    y = torch.cat((torch.ones(100), -torch.ones(400)))
    torch.manual_seed(SEEDS["sel_many_pos"])
    assert orc.compute_indices_selection(y, 60) == G["sel_many_pos"].tolist()
    y2 = torch.cat((torch.ones(7), -torch.ones(400)))
    torch.manual_seed(SEEDS["sel_few_pos"])
    assert orc.compute_indices_selection(y2, 60) == G["sel_few_pos"].tolist()


def test_product_wrapper_centre_selection_matches_reference_rng(tmp_path):
    import FALKONWrapper_with_centers_selection_incore as falkon
    w = falkon.FALKONWrapper(cfg_file(tmp_path))
    torch.manual_seed(SEEDS["sel_many_pos"])
    assert w.compute_indices_selection(torch.cat((torch.ones(100), -torch.ones(400)))) == G["sel_many_pos"].tolist()
    torch.manual_seed(SEEDS["sel_few_pos"])
    assert w.compute_indices_selection(torch.cat((torch.ones(7), -torch.ones(400)))) == G["sel_few_pos"].tolist()


def test_product_feature_statistics_and_normalisation_match_reference(monkeypatch):
    import py_od_utils as UT
    monkeypatch.setattr(UT, "_GPU", "cpu")          # the function ends in .to('cuda') like the reference's; no GPU here
    positives, negatives = inputs()
    torch.manual_seed(SEEDS["stats"])
    stats = UT.computeFeatStatistics_torch(positives, negatives, num_samples=400, features_dim=24, cpu_tensor=True)
    assert torch.equal(stats["mean"].cpu(), t("stats_mean"))
    assert torch.equal(stats["std"].cpu(), t("stats_std"))
    assert torch.equal(stats["mean_norm"].cpu().reshape(1), t("stats_mean_norm"))
    stats_cpu = {k: v.cpu() for k, v in stats.items()}
    COXY = {"C": t("in_reg_C").clone(), "O": None, "X": t("in_reg_X").clone(), "Y": t("in_reg_Y").clone()}
    COXY = UT.normalize_COXY(COXY, stats_cpu, cpu=True)
    assert torch.equal(COXY["X"], t("coxy_X_norm"))
    # OnlineRegionClassifier.zScores == oracle zscores == what trainRegionClassifier left in positives[0]
    z = orc.zscores(positives[0], t("stats_mean"), t("stats_mean_norm")[0])
    assert torch.equal(z.float(), t("zscored_pos0"))


def test_product_shuffle_and_positive_loading_match_reference_rng():
    import py_od_utils as UT
    _, negatives = inputs()
    torch.manual_seed(SEEDS["shuffle"])
    sh = UT.shuffle_negatives(negatives, batch_size=100, num_batches=3)
    assert torch.equal(sh[0][0], t("shuffled_0_0")) and torch.equal(sh[2][2], t("shuffled_2_2"))
    torch.manual_seed(SEEDS["positives_from_coxy"])
    lp = UT.load_positives_from_COXY({"C": t("in_reg_C").clone(), "X": t("in_reg_X").clone()}, samples_fraction=0.5)
    assert torch.equal(lp[1], t("coxy_pos_1"))


def test_oracle_minibootstrap_reproduces_reference_loop():
    """The reference's trainRegionClassifier (z-score, minibootstrap over 4 negative batches, 3 classes, its own RNG
    stream for the centre draws) against the oracle's restatement of the loop: identical caches of surviving
    negatives, identical centres, alpha equal to rounding (both sides solve with the oracle)."""
    positives, negatives = inputs()
    mean, mn = t("stats_mean"), t("stats_mean_norm")[0]
    sigma, lam, M = 12.0, 0.001, 60
    torch.manual_seed(SEEDS["minibootstrap"])

    def train(Xc, y):
        idx = orc.compute_indices_selection(y, M)
        return (Xc[idx], orc.falkon_fit(Xc, y, Xc[idx], sigma, lam, dtype=torch.float64, eps_pc=1e-5, eps_cg=1e-7))

    def predict(model, Xq):
        return orc.falkon_predict(Xq, model[0], model[1].float().double(), sigma).float()

    for i in range(T_CLS):
        pos = orc.zscores(positives[i], mean, mn).float()
        neg = [orc.zscores(b, mean, mn).float() for b in negatives[i]]
        model, cache = orc.minibootstrap(pos, neg, train, predict)
        assert torch.equal(cache, t("cache%d_neg" % i))
        assert torch.equal(model[0], t("model%d_centres" % i))
        assert rel(model[1], t("model%d_alpha" % i)) < 1e-6


def test_oracle_rls_matches_reference_trainer():
    C = t("in_reg_C")[:, 0]
    X = t("coxy_X_norm")
    Y = t("in_reg_Y")
    for i in range(T_CLS):
        sel = C == (i + 1)
        m = orc.rls_train_class(X[sel], Y[sel], CFG["REGION_REFINER"]["opts"]["lambda"])
        W = torch.stack([m["Beta"][str(k)]["weights"] for k in range(4)], 1)
        L = torch.stack([m["Beta"][str(k)]["losses"] for k in range(4)], 1)
        assert rel(m["mu"], t("rls%d_mu" % i)) < 1e-6
        assert rel(m["T"], t("rls%d_T" % i)) < 1e-5 and rel(m["T_inv"], t("rls%d_Tinv" % i)) < 1e-5
        assert rel(W, t("rls%d_W" % i)) < 1e-5
        assert float((L - t("rls%d_losses" % i)).abs().max()) < 1e-6


def test_oracle_box_decode_matches_reference():
    dec = orc.decode_boxes(G["in_test_boxes"], G["in_deltas"], 640, 480)
    # fp32 with one exp per coordinate pair: numpy's and torch's expf may differ in the last bit (1 ulp at 640 = 6e-5)
    assert float(np.abs(dec - G["decoded"]).max()) <= 1.3e-4
    clipped = (G["decoded"] == 0) | (G["decoded"] == 639) | (G["decoded"] == 479)
    assert np.array_equal(dec[clipped], G["decoded"][clipped])


# ------------------------------------------------------------------------------------------ the RPN / segmentation flavours
GF = np.load(os.path.join(HERE, "golden", "reference_flow_flavours.npz"))
METAF = json.load(open(os.path.join(HERE, "golden", "reference_flow_flavours.json")))
FLAVOURS = {"rpn": dict(kw={"is_rpn": True}, pos="rpn_pos", neg="rpn_neg", n=4, batches=3, seed="rpn",
                        line="RPN's Online Classifier training time"),
            "seg": dict(kw={"is_segmentation": True}, pos="seg_pos", neg="seg_neg", n=3, batches=1, seed="segmentation",
                        line="Online Segmentation training time")}


def tf(name):
    return torch.from_numpy(GF[name])


def flavour_inputs(tag):
    f = FLAVOURS[tag]
    pos = [tf("in_%s%d" % (f["pos"], i)).clone() for i in range(f["n"])]
    neg = [[tf("in_%s%d_%d" % (f["neg"], i, j)).clone() for j in range(f["batches"])] for i in range(f["n"])]
    st = "rpn" if tag == "rpn" else "seg"
    stats = {"mean": tf("in_stats_%s_mean" % st), "std": tf("in_stats_%s_std" % st), "mean_norm": tf("in_stats_%s_mean_norm" % st)[0]}
    return pos, neg, stats


@pytest.mark.parametrize("tag", ["rpn", "seg"])
def test_oracle_minibootstrap_reproduces_reference_flavours(tag):
    """The reference's on-line RPN (`is_rpn=True`: RPN section of the config, one model per anchor class, an anchor
    without positives -> None) and on-line segmentation (`is_segmentation=True`: ONLINE_SEGMENTATION section, one
    negative tensor per class in a 1-list -> a single fit, no hard / easy selection) flavours, run from the reference's
    own files (tests/golden/make_reference_golden_flavours.py), against the oracle's restatement of the loop."""
    f, hy = FLAVOURS[tag], METAF["hyper"][tag]
    pos, neg, stats = flavour_inputs(tag)
    assert hy["num_classes"] - 1 == f["n"] == int(GF[tag + "_n_models"][0])
    torch.manual_seed(METAF["seeds"][f["seed"]])

    def train(Xc, y):
        idx = orc.compute_indices_selection(y, hy["M"])
        return (Xc[idx], orc.falkon_fit(Xc, y, Xc[idx], hy["sigma"], hy["lam"], dtype=torch.float64, eps_pc=1e-5, eps_cg=1e-7))

    def predict(model, Xq):
        return orc.falkon_predict(Xq, model[0], model[1].float().double(), hy["sigma"]).float()

    for i in range(f["n"]):
        if len(pos[i]) == 0:
            assert bool(GF[tag + "_is_none"][i])
            continue
        p = orc.zscores(pos[i], stats["mean"], stats["mean_norm"]).float()
        nb = [orc.zscores(b, stats["mean"], stats["mean_norm"]).float() for b in neg[i]]
        model, cache = orc.minibootstrap(p, nb, train, predict, hard_thresh=hy["hard"], easy_thresh=hy["easy"])
        assert torch.equal(cache, tf("%s_cache%d_neg" % (tag, i)))
        assert torch.equal(model[0], tf("%s_model%d_centres" % (tag, i)))
        assert rel(model[1], tf("%s_model%d_alpha" % (tag, i))) < 1e-6
    # how the reference called the third-party package in these flavours: same in-core recipe, the section's sigma / lambda
    ctors = [c for c in METAF["third_party_calls"][tag] if "ctor" in c]
    assert len(ctors) == (f["n"] - int(GF[tag + "_is_none"].sum())) * f["batches"]
    assert all(c["ctor"] == "InCoreFalkon" and c["maxiter"] == 20 and c["sigma"] == hy["sigma"] and c["penalty"] == hy["lam"]
               and c["M"] <= hy["M"] for c in ctors)


class _OracleClassifier:
    """TEST-ONLY classifier for the product's drop-in OnlineRegionClassifier on the CPU: the product wrapper supplies the
    configuration and the centre-selection rule (its own code, the reference's RNG calls), the oracle the arithmetic."""

    def __init__(self, wrapper):
        self.w = wrapper

    def train(self, X, y, sigma=None, lam=None):
        idx = self.w.compute_indices_selection(y)
        idx = [idx] if isinstance(idx, int) else idx
        sigma = self.w.sigma if sigma is None else sigma
        lam = self.w.lam if lam is None else lam
        alpha = orc.falkon_fit(X, y, X[idx], sigma, lam, maxiter=self.w.maxiter, dtype=torch.float64, eps_pc=1e-5, eps_cg=1e-7)
        return {"centres": X[idx].clone(), "alpha": alpha.float(), "sigma": sigma}

    def predict(self, model, X, y=None):
        return orc.falkon_predict(X, model["centres"], model["alpha"].double(), model["sigma"]).float()


@pytest.mark.parametrize("tag", ["rpn", "seg"])
def test_product_region_classifier_host_loop_reproduces_reference_flavours(tag, tmp_path):
    """The PRODUCT's drop-in OnlineRegionClassifier_incore module (configuration sections, class count, in-place
    z-scoring, minibootstrap thresholds, None for an empty class, result.txt line) on the reference's inputs and RNG
    seed, with the oracle standing in for the CUDA fits: same surviving negatives, same centres, same alpha."""
    import FALKONWrapper_with_centers_selection_incore as falkon
    import OnlineRegionClassifier_incore as ocr
    f = FLAVOURS[tag]
    pos, neg, stats = flavour_inputs(tag)
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(yaml.dump(METAF["cfg"]))
    torch.manual_seed(METAF["seeds"][f["seed"]])
    wrapper = falkon.FALKONWrapper(str(cfg), **f["kw"])
    hy = METAF["hyper"][tag]
    assert (wrapper.sigma, wrapper.lam, wrapper.nyst_centers) == (hy["sigma"], hy["lam"], hy["M"])
    rc = ocr.OnlineRegionClassifier(_OracleClassifier(wrapper), pos, neg, stats, cfg_path=str(cfg), **f["kw"])
    assert (rc.num_classes, rc.sigma, rc.lam, rc.hard_tresh, rc.easy_tresh) == (hy["num_classes"], hy["sigma"], hy["lam"], hy["hard"], hy["easy"])
    models, caches = rc.trainRegionClassifier(opts={"return_caches": True}, output_dir=str(tmp_path))
    assert len(models) == f["n"] and [m is None for m in models] == GF[tag + "_is_none"].tolist()
    for i, m in enumerate(models):
        if m is None:
            continue
        assert torch.equal(caches[i]["neg"], tf("%s_cache%d_neg" % (tag, i)))
        assert torch.equal(m["centres"], tf("%s_model%d_centres" % (tag, i)))
        assert rel(m["alpha"], tf("%s_model%d_alpha" % (tag, i))) < 1e-6
    assert open(tmp_path / "result.txt").read().startswith(f["line"]) and METAF["result_lines"][tag].startswith(f["line"])


def test_product_region_classifier_host_loop_reproduces_reference_detection_flavour(tmp_path):
    """Same check for the detection flavour (reference_flow.npz): the product's drop-in module with the oracle-backed
    classifier on the CPU — the GPU test below repeats it with the CUDA fits."""
    import FALKONWrapper_with_centers_selection_incore as falkon
    import OnlineRegionClassifier_incore as ocr
    positives, negatives = inputs()
    stats = {"mean": t("stats_mean"), "std": t("stats_std"), "mean_norm": t("stats_mean_norm")[0]}
    cfg = cfg_file(tmp_path)
    torch.manual_seed(SEEDS["minibootstrap"])
    rc = ocr.OnlineRegionClassifier(_OracleClassifier(falkon.FALKONWrapper(cfg)), positives, negatives, stats, cfg_path=cfg)
    models, caches = rc.trainRegionClassifier(opts={"return_caches": True}, output_dir=str(tmp_path))
    for i in range(T_CLS):
        assert torch.equal(caches[i]["neg"], t("cache%d_neg" % i))
        assert torch.equal(models[i]["centres"], t("model%d_centres" % i))
        assert rel(models[i]["alpha"], t("model%d_alpha" % i)) < 1e-6
    assert torch.equal(positives[0], t("zscored_pos0"))                       # z-scored in place, like the reference
    assert open(tmp_path / "result.txt").read().startswith("Detector's Online Classifier training time")


# ------------------------------------------------------------------------------------------ data formats either side of the path
def test_product_loaders_and_helpers_match_reference_on_the_same_feature_caches(tmp_path, monkeypatch):
    """The feature-cache file formats (positives_cl_{c}_batch_{b}, negatives_cl_{c}_batch_{b}, reg_{x,c,y}_batch_{b}) and
    the helpers around them: the product's drop-in py_od_utils.py against the outputs of the reference's own
    load_features_classifier (plain / cpu_tensor / sample_ratio / shuffled for a `detector` and an `RPN` directory /
    is_segm), load_features_regressor (all rows / a seeded fraction), minibatch_positives, mask_iou and zScores on
    the same files and RNG seeds (tests/golden/make_reference_golden_formats.py)."""
    import sys as _sys
    _sys.path.insert(0, os.path.join(HERE, "golden"))
    import format_fixture as fx
    import py_od_utils as UT
    assert "online-detection_b200" in UT.__file__
    monkeypatch.setattr(UT, "_GPU", "cpu")
    ref = np.load(os.path.join(HERE, "golden", "reference_formats.npz"))
    cfg = fx.build(str(tmp_path))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = fx.run_all(UT, str(tmp_path), cfg)
    got = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}
    assert sorted(got) == sorted(ref.files)
    for k in ref.files:
        assert got[k].shape == ref[k].shape and got[k].dtype == ref[k].dtype, k
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    # spot checks of what the fixture exercises
    assert ref["det_pos1"].shape == (0,) and int(ref["det_n_classes"][0]) == 3          # class without positive files
    assert int(ref["det_shuffled_neg0_n"][0]) == 3 and ref["det_shuffled_neg0_0"].shape[0] == 40
    assert int(ref["rpn_shuffled_neg0_n"][0]) == 2 and ref["rpn_shuffled_neg1_0"].shape[0] == 55
    assert ref["seg_neg1"].shape == (0,) and np.isnan(ref["mask_iou"][4, 3])             # no negatives / empty masks


# ------------------------------------------------------------------------------------------ the detection evaluator (mAP)
def test_oracle_ap_matches_the_reference_evaluator():
    """oracle.detection_ap against the reference's own eval_detection_icw (icw_eval.py:227-402) on synthetic detections
    with duplicates, false positives, a class without ground truth and one without detections: per-class AP (VOC07 and
    area metric, IoU 0.5 and 0.7) and their nanmean, to rounding.  This is the function the "mAP within 0.1" check of
    tests/test_gpu_parity.py is computed with."""
    E = np.load(os.path.join(HERE, "golden", "reference_eval.npz"))
    n = int(E["n_img"][0])
    dets = [(E["in_det_boxes%d" % i], E["in_det_scores%d" % i], E["in_det_labels%d" % i]) for i in range(n)]
    gts = [(E["in_gt_boxes%d" % i], E["in_gt_labels%d" % i]) for i in range(n)]
    for thr in (0.5, 0.7):
        for m07 in (True, False):
            tag = "iou%02d_%s" % (int(thr * 10), "voc07" if m07 else "area")
            ap, m = orc.detection_ap(dets, gts, iou_thresh=thr, use_07_metric=m07)
            ref = E["ap_" + tag]
            assert ap.shape == ref.shape and np.array_equal(np.isnan(ap), np.isnan(ref))
            assert np.allclose(ap, ref, rtol=0, atol=1e-12, equal_nan=True)
            assert abs(m - float(E["map_" + tag][0])) < 1e-12
    ap1, _ = orc.detection_ap(dets[:1], gts[:1])
    assert np.allclose(ap1, E["ap_single"], atol=1e-12, equal_nan=True)
    assert abs(orc.detection_map(dets, gts) - float(E["map_iou05_voc07"][0])) < 1e-12
    # the double +1 of the reference pipeline ([:, 2:] += 1, then boxlist_iou's own +1) is visible: a pair of boxes
    # whose plain IoU is just below 0.5 but matches under the reference's effective widths
    det = [(np.array([[0, 0, 9, 9]], np.float32), np.array([0.9], np.float32), np.array([1]))]
    gt = [(np.array([[0, 3, 9, 12]], np.float32), np.array([1]))]
    plain = (10 * 7) / (2 * 100 - 70)                                  # widths x2 - x1 + 1
    assert plain < 0.55 < (11 * 8) / (2 * 121 - 88)                      # widths x2 - x1 + 2
    assert abs(orc.detection_ap(det, gt, iou_thresh=0.55)[0][1] - 1.0) < 1e-12


# ------------------------------------------------------------------------------------------ the detection post-processor
def test_oracle_postprocessing_matches_the_reference_postprocessor():
    """The reference's OnlineDetectionPostProcessor.forward (decode -> clip -> strict score threshold -> per-class NMS ->
    concatenation in class order -> kthvalue top-K with >=), run from its own file (make_reference_golden_post.py), against
    the oracle's decode_boxes + clip_to_image + filter_results: same detections in the same order, labels and scores
    bit-exact, boxes to fp32 rounding; four threshold / top-K settings incl. "nothing survives".  Also the first-party
    IoU twin compute_overlap_torch against box_iou_plus1."""
    P = np.load(os.path.join(HERE, "golden", "reference_post.npz"))
    dec = orc.clip_to_image(orc.decode_boxes(P["in_proposals"], P["in_deltas"], 640, 480), 640, 480)
    for tag in ("k40", "k100", "tight", "none"):
        thr, nms, k = P[tag + "_params"]
        b, sc, lab, _ = orc.filter_results(dec, P["in_scores"], float(thr), float(nms), int(k))
        assert np.array_equal(lab, P[tag + "_labels"]) and np.array_equal(sc, P[tag + "_scores"]), tag
        assert b.shape == P[tag + "_boxes"].shape and (len(b) == 0 or float(np.abs(b - P[tag + "_boxes"]).max()) <= 1.3e-4), tag
    assert len(P["k40_scores"]) == 40 and len(P["k100_scores"]) == 100 and len(P["none_scores"]) == 0
    iou = orc.box_iou_plus1(P["iou_gt"][None, :], P["in_proposals"])[0]
    twin = P["iou_twin"]
    assert np.allclose(np.where(iou > 0, iou, 0), twin, rtol=0, atol=1e-6) and float(twin.max()) > 0.3 and float(twin.min()) == 0.0


# ------------------------------------------------------------------------------------------ GPU: product modules
@pytest.fixture(scope="module")
def odf():
    import odf as _odf
    return _odf


@pytest.mark.gpu
def test_product_region_classifier_reproduces_reference_flow(odf, tmp_path):
    """Drop-in OnlineRegionClassifier_incore + FALKONWrapper (CUDA fits through the C ABI) on the reference's inputs,
    statistics and RNG seed: the same negatives survive the minibootstrap, the same centres are drawn, the test-time
    score matrix matches within the 1e-3 parity bar."""
    import FALKONWrapper_with_centers_selection_incore as falkon
    import OnlineRegionClassifier_incore as ocr
    positives, negatives = inputs()
    positives = [p.cuda() for p in positives]
    negatives = [[b.cuda() for b in bs] for bs in negatives]
    stats = {"mean": t("stats_mean").cuda(), "std": t("stats_std").cuda(), "mean_norm": t("stats_mean_norm")[0].cuda()}
    cfg = cfg_file(tmp_path)
    torch.manual_seed(SEEDS["minibootstrap"])
    clf = falkon.FALKONWrapper(cfg)
    rc = ocr.OnlineRegionClassifier(clf, positives, negatives, stats, cfg_path=cfg)
    models, caches = rc.trainRegionClassifier(opts={"return_caches": True})
    assert len(models) == T_CLS
    for i in range(T_CLS):
        assert torch.equal(caches[i]["neg"].cpu(), t("cache%d_neg" % i))          # same hard / easy decisions
        assert torch.equal(models[i].ny_points_.cpu(), t("model%d_centres" % i))  # same centre draws
        assert models[i].M == G["model%d_centres" % i].shape[0]
    assert torch.equal(positives[0].cpu(), t("zscored_pos0"))                      # z-scored in place, like the reference
    test = [{"boxes": G["in_test_boxes"], "feat": G["in_test_feat"], "gt": np.zeros(len(G["in_test_boxes"])),
             "img_size": (640, 480)}]
    preds = rc.testRegionClassifier(models, test)
    scores = preds[0].get_field("scores")
    ref = t("test_scores")
    assert scores.shape == ref.shape and torch.equal(scores[:, 0].cpu(), ref[:, 0])   # background column = -1
    assert rel(scores, ref) < 1e-3
    assert torch.equal(scores[:, 1:].argmax(1).cpu(), ref[:, 1:].argmax(1))


@pytest.mark.gpu
def test_product_out_of_core_flavour_reproduces_reference_flow(odf, tmp_path):
    """Drop-in OnlineRegionClassifier + FALKONWrapper_with_centers_selection (features and models parked in host RAM,
    fits staged through HBM by fit()'s side-stream upload): same centres as the reference flow, host-resident models,
    scores within the parity bar."""
    import FALKONWrapper_with_centers_selection as falkon_cpu
    import OnlineRegionClassifier as ocr_cpu
    positives, negatives = inputs()
    stats = {"mean": t("stats_mean"), "std": t("stats_std"), "mean_norm": t("stats_mean_norm")[0]}
    cfg = cfg_file(tmp_path)
    torch.manual_seed(SEEDS["minibootstrap"])
    clf = falkon_cpu.FALKONWrapper(cfg)
    rc = ocr_cpu.OnlineRegionClassifier(clf, positives, negatives, stats, cfg_path=cfg)
    models = rc.trainRegionClassifier()
    x = (t("in_test_feat") - stats["mean"]) * (20 / stats["mean_norm"])
    for i, m in enumerate(models):
        assert not m.ny_points_.is_cuda and not m.alpha_.is_cuda
        assert torch.equal(m.ny_points_, t("ooc_model%d_centres" % i))
        s = clf.predict(m, x)
        assert not s.is_cuda
        ref = orc.falkon_predict(x, t("ooc_model%d_centres" % i), t("ooc_model%d_alpha" % i).double(), 12.0)
        assert rel(s, ref) < 1e-3


@pytest.mark.gpu
def test_product_models_in_the_reference_inference_head(odf, tmp_path):
    """The reference's RoI box head (roi_box_predictors.py:32-160) ran in the golden script with the reference-flow models;
    here the SAME call sequence runs on the product's model objects and regressors: the duck type the head relies on
    (`.M`, `.alpha_`, `.ny_points_`, `.kernel.mmv(features, nystrom_parallel, alpha_parallel)`, `.predict`, truthiness,
    None entries scoring -2) and the RLS apply `x W blockdiag(T_inv) + mu`."""
    import FALKONWrapper_with_centers_selection_incore as falkon
    import OnlineRegionClassifier_incore as ocr
    from region_refiner import RegionRefiner
    positives, negatives = inputs()
    stats = {"mean": t("stats_mean").cuda(), "std": t("stats_std").cuda(), "mean_norm": t("stats_mean_norm")[0].cuda()}
    cfg = cfg_file(tmp_path)
    torch.manual_seed(SEEDS["minibootstrap"])
    clf = falkon.FALKONWrapper(cfg)
    rc = ocr.OnlineRegionClassifier(clf, [p.cuda() for p in positives], [[b.cuda() for b in bs] for bs in negatives], stats,
                                    cfg_path=cfg)
    classifiers = rc.trainRegionClassifier()
    COXY = {"C": t("in_reg_C").cuda(), "O": None, "X": t("coxy_X_norm").cuda(), "Y": t("in_reg_Y").cuda()}
    regressors = RegionRefiner(cfg).trainRegionRefiner(COXY)
    x_raw = t("in_test_feat").cuda()

    # refine_boxes_parallel (:97-124), on the un-normalised features as the head does by default
    W = torch.zeros((x_raw.shape[1] + 1, 4), device="cuda")
    Tinv = torch.zeros(((len(regressors) + 1) * 4,) * 2, device="cuda")
    mu = torch.zeros((1, 4), device="cuda")
    for j, r in enumerate(regressors):
        wj = torch.stack([r["Beta"][str(k)]["weights"] for k in range(4)], 0)
        Tinv[(j + 1) * 4:(j + 2) * 4, (j + 1) * 4:(j + 2) * 4] = r["T_inv"]
        mu = torch.cat((mu, r["mu"].view(1, 4)), 1)
        W = torch.cat((W, wj.t()), 1)
    bbox = (x_raw @ W[:-1] + W[-1]) @ Tinv + mu
    assert rel(bbox, t("head_bbox_par")) < 1e-4

    x = (x_raw - stats["mean"]) * (20 / stats["mean_norm"])                 # forward(): :47-50
    # predict_clss_FALKON_parallel (:140-160)
    assert all(bool(c) for c in classifiers)
    total = sum(c.M for c in classifiers)
    alpha_parallel = torch.zeros((total, len(classifiers)), device="cuda")
    row = 0
    for i, c in enumerate(classifiers):
        alpha_parallel[row:row + c.M, i] = c.alpha_.squeeze()
        row += c.M
    nystrom_parallel = torch.cat([c.ny_points_ for c in classifiers])
    scores = classifiers[0].kernel.mmv(x, nystrom_parallel, alpha_parallel)
    scores = torch.cat((torch.full((x.shape[0], 1), -2.0, device="cuda"), scores), 1)
    ref = t("head_scores_par")
    assert scores.shape == ref.shape and rel(scores, ref) < 1e-3
    assert torch.equal(scores.argmax(1).cpu(), ref.argmax(1))
    # predict_clss_FALKON (:127-138) with a missing classifier
    seq = torch.full((x.shape[0], 1), -2.0, device="cuda")
    for c in (classifiers[0], None, classifiers[2]):
        seq = torch.cat((seq, torch.full((x.shape[0], 1), -2.0, device="cuda") if c is None else c.predict(x)), 1)
    ref = t("head_scores_seq")
    assert torch.equal(seq[:, 2].cpu(), ref[:, 2]) and rel(seq, ref) < 1e-3


@pytest.mark.gpu
def test_product_region_refiner_matches_reference_trainer(odf, tmp_path):
    from region_refiner import RegionRefiner
    COXY = {"C": t("in_reg_C").cuda(), "O": None, "X": t("coxy_X_norm").cuda(), "Y": t("in_reg_Y").cuda()}
    models = RegionRefiner(cfg_file(tmp_path)).trainRegionRefiner(COXY)
    assert len(models) == T_CLS
    for i, m in enumerate(models):
        W = torch.stack([m["Beta"][str(k)]["weights"] for k in range(4)], 1)
        assert rel(m["mu"], t("rls%d_mu" % i)) < 1e-6 and rel(m["T"], t("rls%d_T" % i)) < 1e-5
        assert rel(m["T_inv"], t("rls%d_Tinv" % i)) < 1e-5 and rel(W, t("rls%d_W" % i)) < 1e-5


@pytest.mark.gpu
def test_product_region_predictor_matches_reference(odf, tmp_path):
    """RegionRefiner.predict (predict_regions.py:16-80) with the reference-trained regressors from the fixture: refined
    boxes (n, classes, 4), example boxes in slot 0, np.spacing(1) widths, clipping."""
    from region_refiner import RegionRefiner
    from boxlist import BoxList
    models = np.empty((0,))
    for i in range(T_CLS):
        W = t("rls%d_W" % i).cuda()
        models = np.append(models, {"mu": t("rls%d_mu" % i).cuda(), "T": t("rls%d_T" % i).cuda(), "T_inv": t("rls%d_Tinv" % i).cuda(),
                                    "Beta": {str(k): {"weights": W[:, k].contiguous()} for k in range(4)}})
    rr = RegionRefiner(cfg_file(tmp_path))
    bl = BoxList(t("in_test_boxes").clone(), (640, 480), mode="xyxy")
    feats = [{"feat": G["in_test_feat"], "gt": np.zeros(len(G["in_test_boxes"]))}]
    out = rr.predict([bl], feats, models=models)
    ref = t("refined_boxes")
    assert tuple(out[0].bbox.shape) == tuple(ref.shape)
    assert torch.equal(out[0].bbox[:, 0].cpu(), ref[:, 0])                       # example boxes untouched
    assert float((out[0].bbox.cpu() - ref).abs().max()) < 2e-3                    # fp32 GEMM order + expf, boxes up to 640


@pytest.mark.gpu
def test_product_box_decode_matches_reference(odf):
    import py_od_utils as UT
    from boxlist import BoxList
    bl = BoxList(t("in_test_boxes").cuda(), (640, 480), mode="xyxy")
    dec = UT.decode_boxes_detector(bl, t("in_deltas").cuda())
    assert float((dec.cpu() - t("decoded")).abs().max()) <= 1.3e-4        # expf: last-bit differences only
