"""GPU parity of the index / integer side (include/odf.h: odf_select_indices, odf_gather_rows,
odf_decode_boxes, odf_detect_postprocess) against torch.where semantics and the oracle's restatement of
py_od_utils.decode_boxes_detector (src/py_od_utils.py:247-274) and
OnlineDetectionPostProcessor.filter_results (src/modules/accuracy-evaluator/OnlineDetectionPostProcessor.py:35-79).
Bit-exact for indices and kept boxes; decode within 1e-6 relative (expf)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import falkon_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ops(lib):
    from odf import ops as _ops
    return _ops


@pytest.mark.parametrize("n", [0, 1, 31, 1024, 1025, 5000, 2_500_000])
@pytest.mark.parametrize("strict", [True, False])
def test_select_indices_is_torch_where(ops, n, strict):
    g = torch.Generator().manual_seed(n + 1)
    s = torch.randn(max(n, 1), generator=g)[:n].cuda()
    if n > 10:
        s[3] = -0.7                                    # a value exactly on the threshold
    idx, cnt = ops.select_indices(s, -0.7, strict=strict)
    ref = torch.where(s > -0.7)[0] if strict else torch.where(s >= -0.7)[0]
    k = int(cnt.item())
    assert k == len(ref) and torch.equal(idx[:k], ref)
    # (n, 1) score columns, as FALKONWrapper.predict returns them
    if n:
        idx2, cnt2 = ops.select_indices(s[:, None], -0.7, strict=strict)
        assert int(cnt2.item()) == k and torch.equal(idx2[:k], ref)


def test_select_none_and_all(ops):
    s = torch.linspace(-1, 1, 3000).cuda()
    idx, cnt = ops.select_indices(s, 5.0)
    assert int(cnt.item()) == 0
    idx, cnt = ops.select_indices(s, -5.0)
    assert int(cnt.item()) == 3000 and torch.equal(idx[:3000], torch.arange(3000, device="cuda"))


@pytest.mark.parametrize("d", [1, 7, 256, 1024, 2048])
def test_gather_rows_appends_in_order(ops, d):
    g = torch.Generator().manual_seed(d)
    X = torch.randn(3000, d, generator=g).cuda()
    s = torch.randn(3000, generator=g).cuda()
    idx, cnt = ops.select_indices(s, 0.3)
    dst = torch.full((4000, d), 9.0, device="cuda")
    ops.gather_rows(X, idx, cnt, dst[100:], max_rows=3000)
    k = int(cnt.item())
    ref = X[torch.where(s > 0.3)[0]]
    assert torch.equal(dst[100:100 + k], ref)
    assert bool((dst[:100] == 9).all()) and bool((dst[100 + k:] == 9).all())


def _boxes(rng, R, W=640, H=480):
    xy = np.stack([rng.randint(0, W - 50, R), rng.randint(0, H - 50, R)], 1).astype(np.float32)
    wh = rng.randint(10, 200, size=(R, 2)).astype(np.float32)
    return orc.clip_to_image(np.concatenate([xy, xy + wh], 1), W, H)


def test_decode_boxes_matches_reference_arithmetic(ops):
    rng = np.random.RandomState(0)
    R, Tc = 300, 31
    ex = _boxes(rng, R)
    deltas = (rng.randn(R, 4 * Tc) * 0.2).astype(np.float32)
    ref = orc.decode_boxes(ex, deltas, 640, 480)
    out = ops.decode_boxes(torch.from_numpy(ex).cuda(), torch.from_numpy(deltas).cuda(), 640, 480).cpu().numpy()
    assert np.abs(out - ref).max() <= 1e-3 and np.abs(out - ref).max() / np.abs(ref).max() < 1e-6
    assert out[:, 0::4].min() >= 0 and out[:, 2::4].max() <= 639 and out[:, 3::4].max() <= 479


@pytest.mark.parametrize("R,Tc,K,seed", [(300, 31, 100, 0), (300, 22, 100, 1), (1000, 16, 300, 2), (64, 2, 100, 3),
                                          (1, 5, 100, 4), (2100, 3, 50, 5)])
def test_detect_postprocess_bit_identical_to_filter_results(ops, R, Tc, K, seed):
    rng = np.random.RandomState(seed)
    box = _boxes(rng, R)
    boxes = np.tile(box, (1, Tc)) + rng.randint(-3, 4, size=(R, 4 * Tc)).astype(np.float32)
    boxes = np.concatenate([orc.clip_to_image(boxes[:, 4 * j:4 * j + 4], 640, 480) for j in range(Tc)], 1)
    scores = (rng.rand(R, Tc).astype(np.float32) * 3 - 2.2)            # about 7 % below the -2 threshold
    scores[:, 0] = -1
    rb, rs, rl, keeps = orc.filter_results(boxes, scores, -2.0, 0.3, K)
    ob, os_, ol, orr = ops.detect_postprocess(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), -2.0, 0.3, K)
    assert len(os_) == len(rs)
    assert np.array_equal(ob.cpu().numpy(), rb) and np.array_equal(os_.cpu().numpy(), rs) and np.array_equal(ol.cpu().numpy(), rl)
    # source RoIs: the per-class keep indices of the oracle, concatenated, then the same top-K mask
    all_keep = np.concatenate(keeps) if keeps else np.zeros(0, np.int64)
    all_sc = np.concatenate([scores[k, j + 1] for j, k in enumerate(keeps)]) if keeps else np.zeros(0, np.float32)
    if len(all_sc) > K > 0:
        kth = np.sort(all_sc)[len(all_sc) - K]
        all_keep = all_keep[all_sc >= kth]
    assert np.array_equal(orr.cpu().numpy(), all_keep)


def test_detect_postprocess_edges(ops):
    # nothing above the threshold
    b = torch.zeros(10, 8).cuda()
    s = torch.full((10, 2), -3.0).cuda()
    ob, os_, ol, orr = ops.detect_postprocess(b, s, -2.0, 0.3, 100)
    assert len(os_) == 0
    # identical boxes: only the best survives; score ties are broken by the lower index (stable order)
    box = torch.tensor([[10., 10., 50., 50.]]).repeat(6, 1)
    boxes = torch.cat([box, box], 1).cuda()
    sc = torch.tensor([[-1, .5], [-1, .9], [-1, .9], [-1, .1], [-1, -5.], [-1, .2]]).cuda()
    ob, os_, ol, orr = ops.detect_postprocess(boxes, sc, -2.0, 0.3, 100)
    assert orr.tolist() == [1] and ol.tolist() == [1]
    # top-K with ties at the K-th score keeps all of them (kthvalue rule)
    far = torch.tensor([[0., 0., 5., 5.], [100., 100., 105., 105.], [200., 200., 205., 205.], [300., 300., 305., 305.]])
    boxes = torch.cat([far, far], 1).cuda()
    sc = torch.tensor([[-1, .3], [-1, .7], [-1, .3], [-1, .3]]).cuda()
    ob, os_, ol, orr = ops.detect_postprocess(boxes, sc, -2.0, 0.3, 2)
    assert orr.tolist() == [0, 1, 2, 3]
    rb, rs, rl, _ = orc.filter_results(boxes.cpu().numpy(), sc.cpu().numpy(), -2.0, 0.3, 2)
    assert np.array_equal(os_.cpu().numpy(), rs)


def test_postprocessor_dropin_module(ops):
    sys.path.insert(0, os.path.join(ROOT, "online-detection_b200", "modules", "accuracy-evaluator"))
    sys.path.insert(0, os.path.join(ROOT, "online-detection_b200", "modules"))
    from OnlineDetectionPostProcessor import OnlineDetectionPostProcessor
    from boxlist import BoxList
    rng = np.random.RandomState(7)
    R, Tc = 300, 6
    ex = _boxes(rng, R)
    deltas = (rng.randn(R, 4 * Tc) * 0.1).astype(np.float32)
    scores = (rng.rand(R, Tc).astype(np.float32) * 3 - 2.2)
    scores[:, 0] = -2
    pp = OnlineDetectionPostProcessor(score_thresh=-2, nms=0.3, detections_per_img=100)
    props = [BoxList(torch.from_numpy(ex), (640, 480))]
    res = pp((torch.from_numpy(scores), torch.from_numpy(deltas)), props, Tc, (640, 480))
    dec = orc.clip_to_image_multi(orc.decode_boxes(ex, deltas, 640, 480), 640, 480) if hasattr(orc, "clip_to_image_multi") \
        else orc.decode_boxes(ex, deltas, 640, 480)
    # run the oracle on the GPU-decoded boxes (decode differs from numpy only by expf rounding)
    dec_gpu = ops.decode_boxes(torch.from_numpy(ex).cuda(), torch.from_numpy(deltas).cuda(), 640, 480).cpu().numpy()
    assert np.abs(dec_gpu - dec).max() < 1e-3
    rb, rs, rl, _ = orc.filter_results(dec_gpu, scores, -2.0, 0.3, 100)
    assert np.array_equal(res.bbox.cpu().numpy(), rb)
    assert np.array_equal(res.get_field("scores").cpu().numpy(), rs)
    assert np.array_equal(res.get_field("labels").cpu().numpy(), rl)
