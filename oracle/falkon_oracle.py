"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this file.
It restates, in plain PyTorch-CPU (fp64 by default, fp32 on request), the algorithm that the
reference's on-line learners run through the third-party `falkon` package:

    reference: hsp-iit/online-detection, src/modules/region-classifier/
               FALKONWrapper_with_centers_selection_incore.py:43-82  (train / predict)
               FALKONWrapper_with_centers_selection.py:42-78          (out-of-core flavour)
    arithmetic: github.com/FalkonML/falkon @ 0d96c685dbdff7048e7410e5ca419b21e337789d
               (INSTALLATION_GUIDE.md:67-77) — NOT vendored under /root/reference and not
               installable offline, so its published algorithm (Rudi, Carratino, Rosasco 2017;
               Meanti et al. 2020) is restated here following SURVEY.md Appendix A.

PARITY UNPINNED for the FALKON arithmetic (Gaussian kernel products, preconditioner, CG): the reference
ships no tests, golden vectors or fixtures for this path and the `falkon` package cannot be imported
here, so that part of the oracle is pinned only by (i) the Nystrom kernel-ridge solution it must converge
to -- in closed form and as computed by an independent library (scikit-learn Nystroem + Ridge) --,
(ii) fp64-vs-fp32 self-agreement and (iii) hand-computed known-answer cases for the integer
post-processing (tests/test_oracle.py).

PINNED against the reference's own first-party code (tests/golden/make_reference_golden.py runs the
UNMODIFIED py_od_utils.py, FALKONWrapper_with_centers_selection_incore.py, MyCenterSelector.py,
OnlineRegionClassifier_incore.py and region_refiner.py + trainer on the CPU, with this oracle standing
in for the absent `falkon` package; tests/test_reference_golden.py): centre selection with the
reference's RNG draws, z-scoring, the minibootstrap loop (surviving negatives and centres bit-exact),
feature statistics, COXY normalisation, RLS refiners (mu, T, T_inv, weights, losses), box decode; the
on-line RPN / segmentation flavours of the loop (make_reference_golden_flavours.py); the detection
post-processor and the first-party IoU twin (make_reference_golden_post.py); the VOC-style evaluator
behind the mAP criterion (make_reference_golden_eval.py); the feature-cache loaders are pinned for the
product's py_od_utils directly (make_reference_golden_formats.py).
"""
import math

import numpy as np
import torch

# ------------------------------------------------------------------------------------------
# Gaussian kernel and the three kernel products
# ------------------------------------------------------------------------------------------


def gaussian_kernel(X1, X2, sigma, dtype=torch.float64):
    """K[i, j] = exp(-|x1_i - x2_j|^2 / (2 sigma^2)), evaluated in falkon's order of operations:
    D = X1 X2^T; D *= -2; D += |x1|^2 + |x2|^2; clamp >= 0; D *= -1/(2 sigma^2); exp."""
    X1 = X1.to(dtype)
    X2 = X2.to(dtype)
    D = X1 @ X2.T
    D *= -2.0
    D += (X1 * X1).sum(1)[:, None]
    D += (X2 * X2).sum(1)[None, :]
    D.clamp_(min=0)
    D *= -1.0 / (2.0 * float(sigma) ** 2)
    return D.exp_()


def _row_blocks(n, block):
    for s in range(0, n, block):
        yield s, min(n, s + block)


def mmv(X1, X2, v, sigma, dtype=torch.float64, block=4096):
    """GaussianKernel.mmv: K(X1, X2) @ v, blocked over rows of X1 (falkon fmmv)."""
    v = v.to(dtype)
    out = torch.empty(X1.shape[0], v.shape[1], dtype=dtype)
    for s, e in _row_blocks(X1.shape[0], block):
        out[s:e] = gaussian_kernel(X1[s:e], X2, sigma, dtype) @ v
    return out


def dmmv(X1, X2, v, w, sigma, dtype=torch.float64, block=4096):
    """GaussianKernel.dmmv: K(X1,X2)^T (K(X1,X2) v + w); v or w may be None (falkon fdmmv)."""
    T = (v if v is not None else w).shape[1]
    out = torch.zeros(X2.shape[0], T, dtype=dtype)
    for s, e in _row_blocks(X1.shape[0], block):
        Kb = gaussian_kernel(X1[s:e], X2, sigma, dtype)
        inner = torch.zeros(e - s, T, dtype=dtype)
        if v is not None:
            inner += Kb @ v.to(dtype)
        if w is not None:
            inner += w[s:e].to(dtype)
        out += Kb.T @ inner
    return out


# ------------------------------------------------------------------------------------------
# Preconditioner (Appendix A.3) and preconditioned CG (Appendix A.4)
# ------------------------------------------------------------------------------------------


def pc_epsilon(dtype):
    return 1e-5 if dtype == torch.float32 else 1e-13


def cg_epsilon(dtype):
    return 1e-7 if dtype == torch.float32 else 1e-15


class Preconditioner:
    """T = chol_upper(K_MM + eps*M*I) (K ~ T^T T);  A = chol_upper(T T^T / M + lam I)."""

    def __init__(self, C, sigma, lam, dtype=torch.float64, eps=None, kmm_dtype=torch.float64):
        M = C.shape[0]
        eps = pc_epsilon(dtype) if eps is None else eps
        # falkon evaluates K_MM in double precision when data is single (no_single_kernel=True)
        K = gaussian_kernel(C, C, sigma, kmm_dtype).to(dtype)
        K += eps * M * torch.eye(M, dtype=dtype)
        self.T = torch.linalg.cholesky(K, upper=True)
        G = self.T @ self.T.T / M
        G += lam * torch.eye(M, dtype=dtype)
        self.A = torch.linalg.cholesky(G, upper=True)

    def invT(self, v):
        return torch.linalg.solve_triangular(self.T, v, upper=True)

    def invTt(self, v):
        return torch.linalg.solve_triangular(self.T.T, v, upper=False)

    def invA(self, v):
        return torch.linalg.solve_triangular(self.A, v, upper=True)

    def invAt(self, v):
        return torch.linalg.solve_triangular(self.A.T, v, upper=False)

    def apply(self, v):
        return self.invT(self.invA(v))

    def apply_t(self, v):
        return self.invAt(self.invTt(v))


def falkon_fit(X, Y, centres, sigma, lam, maxiter=20, dtype=torch.float64, tol=1e-7,
               full_gradient_every=10, eps_pc=None, eps_cg=None, return_trace=False, cache_knm=False):
    """InCoreFalkon.fit restated (A.4).  X (N x d), Y (N x T) or (N,), centres (M x d) already
    selected (MyCenterSelector.select == X[indices]).  Returns alpha (M x T).

    cache_knm: evaluate K_NM once and reuse it in all 23 sweeps, as upstream does for the reference's sizes
    (`store_kernel_d_threshold=250`, FALKONWrapper_with_centers_selection_incore.py:56) -- the same
    arithmetic up to the blocking of the row sums; the timed CPU baselines of bench.py use it so that the port is not
    a straw man."""
    X = X.to(dtype)
    C = centres.to(dtype)
    Y = Y.to(dtype)
    if Y.dim() == 1:
        Y = Y[:, None]
    N = X.shape[0]
    eps = cg_epsilon(dtype) if eps_cg is None else eps_cg
    pc = Preconditioner(C, sigma, lam, dtype, eps=eps_pc)

    if cache_knm:
        Knm = torch.empty((N, C.shape[0]), dtype=dtype)
        for s_, e_ in _row_blocks(N, 8192):
            Knm[s_:e_] = gaussian_kernel(X[s_:e_], C, sigma, dtype)

        def sweep(v, w):
            inner = Knm @ v if v is not None else None
            if w is not None:
                inner = w if inner is None else inner + w
            return Knm.T @ inner
    else:
        def sweep(v, w):
            return dmmv(X, C, v, w, sigma, dtype)

    B = pc.apply_t(sweep(None, Y / N))

    def op(s):
        v = pc.invA(s)
        u = pc.invT(v)
        c = sweep(u, None) / N
        return pc.invAt(pc.invTt(c) + lam * v)

    beta = torch.zeros_like(B)
    R = B.clone()
    P = R.clone()
    rs_old = (R * R).sum(0)
    trace = []
    for i in range(maxiter):
        AP = op(P)
        a = rs_old / ((P * AP).sum(0) + eps)
        beta += P * a
        if (i + 1) % full_gradient_every == 0:
            R = B - op(beta)
        else:
            R -= AP * a
        rs_new = (R * R).sum(0)
        if return_trace:
            trace.append(rs_new.clone())
        if math.sqrt(float(rs_new.abs().max())) < tol:
            break
        P = R + P * (rs_new / (rs_old + eps))
        rs_old = rs_new
    alpha = pc.apply(beta)
    if return_trace:
        return alpha, trace
    return alpha


def falkon_predict(Xt, centres, alpha, sigma, dtype=torch.float64):
    """Falkon.predict == kernel.mmv(X, ny_points_, alpha_)."""
    return mmv(Xt.to(dtype), centres.to(dtype), alpha.to(dtype), sigma, dtype)


def nystrom_krr_closed_form(X, Y, centres, sigma, lam, eps_pc=0.0):
    """alpha* = (K_nm^T K_nm + lam N (K_mm + eps M I))^-1 K_nm^T Y  in fp64 (pin for the CG)."""
    dt = torch.float64
    X, C, Y = X.to(dt), centres.to(dt), Y.to(dt)
    if Y.dim() == 1:
        Y = Y[:, None]
    N, M = X.shape[0], C.shape[0]
    Knm = gaussian_kernel(X, C, sigma, dt)
    Kmm = gaussian_kernel(C, C, sigma, dt) + eps_pc * M * torch.eye(M, dtype=dt)
    H = Knm.T @ Knm + lam * N * Kmm
    return torch.linalg.lstsq(H, Knm.T @ Y).solution


# ------------------------------------------------------------------------------------------
# Wrapper-level semantics (first-party reference code)
# ------------------------------------------------------------------------------------------


def compute_indices_selection(y, M, generator=None):
    """FALKONWrapper.compute_indices_selection (…incore.py:87-99): at most M/2 positive centres,
    the rest negatives, both sampled WITH replacement (torch.randint) when over budget."""
    pos = (y == 1).nonzero()
    if pos.size(0) > int(M / 2):
        pos = pos[torch.randint(pos.size(0), (int(M / 2),), generator=generator)]
    neg = (y == -1).nonzero()
    if neg.size(0) > M - pos.size(0):
        neg = neg[torch.randint(neg.size(0), (M - pos.size(0),), generator=generator)]
    idx = torch.cat((pos, neg), dim=0).squeeze().tolist()
    if isinstance(idx, int):
        idx = [idx]
    return idx


def zscores(feat, mean, mean_norm, target_norm=20.0):
    """OnlineRegionClassifier.zScores (OnlineRegionClassifier.py:224-227)."""
    return (feat - mean) * (target_norm / float(mean_norm))


def feat_statistics(X_sample):
    """Arithmetic of computeFeatStatistics_torch (py_od_utils.py:88-93) on an already sampled
    matrix: column mean, unbiased column std, mean of the (uncentred) row norms."""
    return X_sample.mean(0), X_sample.std(0), X_sample.norm(dim=1).mean()


def minibootstrap(positives, negative_batches, train_fn, predict_fn, hard_thresh=-0.7, easy_thresh=-0.9):
    """trainWithMinibootstrap for one class (OnlineRegionClassifier_incore.py:99-140)."""
    cache_neg = negative_batches[0]
    model = None
    for j in range(len(negative_batches)):
        if j > 0:
            pred = predict_fn(model, negative_batches[j])
            hard = torch.where(pred > hard_thresh)[0]
            cache_neg = torch.cat((cache_neg, negative_batches[j][hard]), 0)
        X = torch.cat((positives, cache_neg), 0)
        y = torch.cat((torch.ones(len(positives)), -torch.ones(len(cache_neg))), 0)
        model = train_fn(X, y)
        if len(cache_neg) != 0 and j != len(negative_batches) - 1:
            pred = predict_fn(model, cache_neg)
            keep = torch.where(pred >= easy_thresh)[0]
            cache_neg = cache_neg[keep]
    return model, cache_neg


# ------------------------------------------------------------------------------------------
# RLS box refiners (region-refiner/region_refiner_trainer/train_region_refiner.py:25-119)
# ------------------------------------------------------------------------------------------


def rls_train_class(X, Y, lam):
    """One class of RegionRefinerTrainer.train (:54-71) + solve (:100-119), fp64.
    torch.eig (removed in torch 2) -> eigh: S is symmetric 4x4 and T, T_inv do not depend on the
    eigenvector order or sign."""
    Xi = X.to(torch.float64)
    Yi = Y.to(torch.float64).clone()
    Xi = torch.cat((Xi, torch.ones(Xi.shape[0], 1, dtype=torch.float64)), 1)
    mu = Yi.mean(0)
    Yi -= mu
    S = Yi.T @ Yi / Yi.shape[0]
    D, W = torch.linalg.eigh(S)
    T = W @ torch.diag(torch.sqrt(D + 0.001).pow(-1)) @ W.T
    T_inv = W @ torch.diag(torch.sqrt(D + 0.001)) @ W.T
    Yi = Yi @ T
    G = Xi.T @ Xi + lam * torch.eye(Xi.shape[1], dtype=torch.float64)
    R = torch.linalg.cholesky(G)
    beta = {}
    for k in range(4):
        z = torch.linalg.solve_triangular(R, (Xi.T @ Yi[:, k])[:, None], upper=False)
        w = torch.linalg.solve_triangular(R.T, z, upper=True)[:, 0]
        losses = 0.5 * (Xi @ w - Yi[:, k]) ** 2
        beta[str(k)] = {"weights": w.float(), "losses": losses.float()}
    return {"mu": mu.float(), "T": T.float(), "T_inv": T_inv.float(), "Beta": beta}


# ------------------------------------------------------------------------------------------
# Integer post-processing, legacy "+1" box convention
# ------------------------------------------------------------------------------------------


def decode_boxes(ex_boxes, deltas, img_w, img_h):
    """py_od_utils.decode_boxes_detector (:247-274) in numpy float32."""
    ex = np.asarray(ex_boxes, dtype=np.float32)
    d = np.asarray(deltas, dtype=np.float32)
    w = ex[:, 2] - ex[:, 0] + 1
    h = ex[:, 3] - ex[:, 1] + 1
    cx = ex[:, 0] + np.float32(0.5) * w
    cy = ex[:, 1] + np.float32(0.5) * h
    pcx = d[:, 0::4] * w[:, None] + cx[:, None]
    pcy = d[:, 1::4] * h[:, None] + cy[:, None]
    pw = np.exp(d[:, 2::4]) * w[:, None]
    ph = np.exp(d[:, 3::4]) * h[:, None]
    out = np.zeros_like(d)
    out[:, 0::4] = np.maximum(pcx - np.float32(0.5) * pw, 0)
    out[:, 1::4] = np.maximum(pcy - np.float32(0.5) * ph, 0)
    out[:, 2::4] = np.minimum(pcx + np.float32(0.5) * pw - 1, img_w - 1)
    out[:, 3::4] = np.minimum(pcy + np.float32(0.5) * ph - 1, img_h - 1)
    return out


def clip_to_image(boxes, img_w, img_h):
    """BoxList.clip_to_image(remove_empty=False), TO_REMOVE = 1."""
    b = np.array(boxes, dtype=np.float32, copy=True)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, img_w - 1)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, img_h - 1)
    return b


def box_iou_plus1(a, b):
    """boxlist_iou / compute_overlap_torch (mrcnn_modified/utils/evaluations.py:4-18)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    area_a = (a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)
    area_b = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    lt = np.maximum(a[:, None, :2], b[None, :, :2])
    rb = np.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = np.clip(rb - lt + 1, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter)


def nms_plus1(boxes, scores, thresh):
    """maskrcnn-benchmark `_C.nms` semantics: descending score order, suppress when IoU(+1) >
    thresh (strict); returns kept ORIGINAL indices sorted ascending."""
    boxes = np.asarray(boxes, dtype=np.float32)
    order = np.argsort(-np.asarray(scores, dtype=np.float32), kind="stable")
    suppressed = np.zeros(len(order), dtype=bool)
    keep = []
    iou = box_iou_plus1(boxes, boxes) if len(boxes) else np.zeros((0, 0), np.float32)
    for ii, i in enumerate(order):
        if suppressed[i]:
            continue
        keep.append(int(i))
        for j in order[ii + 1:]:
            if not suppressed[j] and iou[i, j] > thresh:
                suppressed[j] = True
    return np.array(sorted(keep), dtype=np.int64)


def filter_results(boxes, scores, score_thresh=-2.0, nms_thresh=0.3, dets_per_img=100):
    """OnlineDetectionPostProcessor.filter_results (AE/OnlineDetectionPostProcessor.py:35-79).
    boxes (R, 4*num_classes), scores (R, num_classes); class 0 is background.
    Returns (boxes[k,4], scores[k], labels[k], keep_per_class: list of index arrays)."""
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    num_classes = scores.shape[1]
    out_b, out_s, out_l, keeps = [], [], [], []
    for j in range(1, num_classes):
        inds = np.nonzero(scores[:, j] > score_thresh)[0]
        sj = scores[inds, j]
        bj = boxes[inds, 4 * j:4 * (j + 1)]
        k = nms_plus1(bj, sj, nms_thresh)
        keeps.append(inds[k])
        out_b.append(bj[k])
        out_s.append(sj[k])
        out_l.append(np.full(len(k), j, dtype=np.int64))
    out_b = np.concatenate(out_b) if out_b else np.zeros((0, 4), np.float32)
    out_s = np.concatenate(out_s) if out_s else np.zeros((0,), np.float32)
    out_l = np.concatenate(out_l) if out_l else np.zeros((0,), np.int64)
    n = len(out_s)
    if n > dets_per_img > 0:
        kth = np.sort(out_s)[n - dets_per_img]      # kthvalue(scores, n - K + 1)
        sel = out_s >= kth
        out_b, out_s, out_l = out_b[sel], out_s[sel], out_l[sel]
    return out_b, out_s, out_l, keeps


def voc07_ap(rec, prec):
    """11-point VOC07 AP (icw_eval.py calc_detection_icw_ap, use_07_metric=True)."""
    ap = 0.0
    for t in np.arange(0.0, 1.1, 0.1):
        p = np.max(np.nan_to_num(prec)[rec >= t]) if np.sum(rec >= t) > 0 else 0.0
        ap += p / 11.0
    return ap


def voc_area_ap(rec, prec):
    """Area under the monotone precision envelope (icw_eval.py:383-400, use_07_metric=False)."""
    mpre = np.concatenate(([0], np.nan_to_num(prec), [0]))
    mrec = np.concatenate(([0], rec, [1]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def detection_ap(dets, gts, iou_thresh=0.5, use_07_metric=True):
    """eval_detection_icw restated (mrcnn_modified/data/datasets/evaluation/icubworld/icw_eval.py:227-402, no
    `difficult` objects): per-class AP array (index = label, nan where a class has no ground truth) and their nanmean.
    dets: list per image of (boxes[k,4], scores[k], labels[k]); gts: list per image of (boxes[g,4], labels[g]).

    As in the reference, predicted and ground-truth boxes get `[:, 2:] += 1` (:289-292) and are THEN compared with
    maskrcnn-benchmark's `boxlist_iou`, which applies its own +1 convention to areas and intersections (SURVEY
    Appendix B) — the effective widths are x2 - x1 + 2.  Per image and label the predictions are walked in
    `argsort()[::-1]` order (:259-262), a ground-truth box is matched at most once (:302-314), and the global ranking is
    again `argsort()[::-1]` of the concatenated scores (:331-332); precision is tp / (tp + fp) with nan for 0 / 0."""
    n_pos, score, match = {}, {}, {}
    for (db, ds, dl), (gb, gl) in zip(dets, gts):
        db, ds, dl = np.asarray(db, dtype=np.float32).reshape(-1, 4), np.asarray(ds), np.asarray(dl)
        gb, gl = np.asarray(gb, dtype=np.float32).reshape(-1, 4), np.asarray(gl)
        for l in np.unique(np.concatenate((dl, gl)).astype(int)):
            b, sc = db[dl == l], ds[dl == l]
            order = sc.argsort()[::-1]
            b, sc = b[order], sc[order]
            g = gb[gl == l]
            n_pos[l] = n_pos.get(l, 0) + len(g)
            score.setdefault(l, []).extend(sc.tolist())
            match.setdefault(l, [])
            if len(b) == 0:
                continue
            if len(g) == 0:
                match[l].extend([0] * len(b))
                continue
            bb, gg = b.copy(), g.copy()
            bb[:, 2:] += 1
            gg[:, 2:] += 1
            iou = box_iou_plus1(bb, gg)
            gidx = iou.argmax(axis=1)
            gidx[iou.max(axis=1) < iou_thresh] = -1
            used = np.zeros(len(g), dtype=bool)
            for gi in gidx:
                if gi >= 0:
                    match[l].append(0 if used[gi] else 1)
                    used[gi] = True
                else:
                    match[l].append(0)
    if not n_pos:
        return np.array([np.nan]), float("nan")
    ap = np.full(max(n_pos) + 1, np.nan)
    for l in n_pos:
        sc, mt = np.array(score[l]), np.array(match[l], dtype=np.int8)
        mt = mt[sc.argsort()[::-1]]
        tp, fp = np.cumsum(mt == 1), np.cumsum(mt == 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            prec = tp / (fp + tp)
        if n_pos[l] > 0:
            rec = tp / n_pos[l]
            ap[l] = voc07_ap(rec, prec) if use_07_metric else voc_area_ap(rec, prec)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = float(np.nanmean(ap))
    return ap, m


def detection_map(dets, gts, num_classes=None, iou_thresh=0.5):
    """VOC07 mAP as the reference's evaluator reports it (`detection_ap`, USE_VOC07_METRIC: True); 0 when no class has
    ground truth.  `num_classes` is accepted for the older call sites and not needed."""
    _ap, m = detection_ap(dets, gts, iou_thresh=iou_thresh, use_07_metric=True)
    return 0.0 if np.isnan(m) else m


# ------------------------------------------------------------------------------------------
# Synthetic workload shared by tests and bench
# ------------------------------------------------------------------------------------------


def make_synthetic(N, d, T, seed=0, pos_fraction=0.1, noise=0.7, dtype=torch.float32, proto_seed=0):
    """Class prototypes mu_t ~ N(0, I) (drawn from `proto_seed`, so train and test sets of one
    problem share them); 10 % positives spread over T classes, 90 % background;
    x = mu_c + noise * N(0, I); then mean-centred and scaled to mean row norm 20.
    Returns X (N x d), labels c (N,) in 0..T (0 = background), Y (N x T) in {+1, -1}."""
    protos = torch.randn(T + 1, d, generator=torch.Generator().manual_seed(7919 + proto_seed))
    g = torch.Generator().manual_seed(seed)
    c = torch.zeros(N, dtype=torch.int64)
    n_pos = int(N * pos_fraction)
    pos_idx = torch.randperm(N, generator=g)[:n_pos]
    c[pos_idx] = torch.randint(1, T + 1, (n_pos,), generator=g)
    X = protos[c] + noise * torch.randn(N, d, generator=g)
    X = X - X.mean(0)
    X = X * (20.0 / X.norm(dim=1).mean())
    Y = -torch.ones(N, T)
    Y[torch.arange(N)[c > 0], c[c > 0] - 1] = 1.0
    return X.to(dtype), c, Y.to(dtype)


def shared_centres(c, M, seed=1):
    """Batched-mode centre set: <= M/2 positives (any class), rest background, without
    replacement where possible; deterministic."""
    g = torch.Generator().manual_seed(seed)
    pos = (c > 0).nonzero()[:, 0]
    neg = (c == 0).nonzero()[:, 0]
    n_pos = min(len(pos), M // 2)
    pos = pos[torch.randperm(len(pos), generator=g)[:n_pos]]
    n_neg = min(len(neg), M - n_pos)
    neg = neg[torch.randperm(len(neg), generator=g)[:n_neg]]
    return torch.cat((pos, neg))
