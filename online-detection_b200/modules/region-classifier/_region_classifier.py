"""Shared implementation of OnlineRegionClassifier (both flavours).

Reference: src/modules/region-classifier/OnlineRegionClassifier_incore.py:16-224 and
OnlineRegionClassifier.py:19-227.  Same constructor, options, return values and side effects
(positives/negatives are overwritten in place with their z-scored versions; a timing line is
appended to <output_dir>/result.txt).  One binary FALKON model per class, trained by
minibootstrap: every negative batch after the first contributes only its hard negatives
(score > HARD_THRESH) and after every refit the cache sheds its easy ones (score < EASY_THRESH).
"""
import os
import sys
import time

import numpy as np
import torch
import yaml

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir)))
import _paths  # noqa: E402,F401
import RegionClassifierAbstract as rcA  # noqa: E402

try:  # the real thing if the CNN stack is installed next to us, else the stand-in
    from maskrcnn_benchmark.structures.bounding_box import BoxList
except Exception:  # noqa: BLE001
    from boxlist import BoxList


class OnlineRegionClassifierBase(rcA.RegionClassifierAbstract):
    HOST_CACHE = False      # "--CPU" flavour keeps the per-class caches in host RAM

    def __init__(self, classifier, positives, negatives, stats=None, cfg_path=None, is_rpn=False,
                 is_segmentation=False):
        if cfg_path is not None:
            with open(cfg_path) as fh:
                self.cfg = yaml.load(fh, Loader=yaml.FullLoader)
            if is_rpn:
                self.cfg = self.cfg["RPN"]
            section = self.cfg["ONLINE_SEGMENTATION" if is_segmentation else "ONLINE_REGION_CLASSIFIER"]
            self.classifier_options = section["CLASSIFIER"]
            self.lam = section["CLASSIFIER"]["lambda"]
            self.sigma = section["CLASSIFIER"]["sigma"]
            self.hard_tresh = section["MINIBOOTSTRAP"]["HARD_THRESH"]
            self.easy_tresh = section["MINIBOOTSTRAP"]["EASY_THRESH"]
            self.mean = 0
            self.std = 0
            self.mean_norm = 0
            self.is_rpn = is_rpn
        else:
            print("Config file path not given. cfg variable set to None.")
            self.cfg = None
        self.classifier = classifier
        self.negatives = negatives
        self.positives = positives
        self.num_classes = len(self.cfg["CHOSEN_CLASSES"]) + (1 if is_rpn else 0)
        if stats:
            self.stats = stats
            self.mean = stats["mean"]
            self.std = stats["std"]
            self.mean_norm = stats["mean_norm"]
        self.normalized = False
        self.is_segmentation = is_segmentation
        self.return_caches = False

    def loadRegionClassifier(self):
        pass

    def processOptions(self, opts):
        for key, attr in (("num_classes", "num_classes"), ("imset_train", "train_imset"),
                          ("classifier_options", "classifier_options"), ("is_rpn", "is_rpn"), ("lam", "lam"),
                          ("sigma", "sigma"), ("return_caches", "return_caches"), ("normalized", "normalized")):
            if key in opts:
                setattr(self, attr, opts[key])

    # ------------------------------------------------------------------ one refit
    def updateModel(self, cache):
        X_pos, X_neg = cache["pos"], cache["neg"]
        X = torch.cat((X_pos, X_neg), 0)
        y = torch.cat((torch.ones(len(X_pos), device=X.device), -torch.ones(len(X_neg), device=X.device)), 0)
        if self.sigma is not None and self.lam is not None:
            print("Updating model with lambda: {} and sigma: {}".format(self.lam, self.sigma))
            return self.classifier.train(X, y, sigma=self.sigma, lam=self.lam)
        print("Updating model with default lambda and sigma")
        return self.classifier.train(X, y)

    def _stage(self, t):
        return t.cpu() if self.HOST_CACHE else t

    # ------------------------------------------------------------------ GPU-resident cache (in-core flavour)
    # The reference grows / prunes the per-class cache with torch.where + indexing + torch.cat
    # (…_incore.py:117-119, 133-135), i.e. a fresh allocation and copy of the whole cache per step.  Here the
    # cache is ONE pre-allocated buffer [positives | negatives] (capacity = all batches of the class) with a
    # twin for pruning; hard / easy rows are picked by odf_select_indices (stable compaction, identical to
    # torch.where) and moved by odf_gather_rows straight into place; X = buffer[:P + n] needs no torch.cat.
    def _gpu_cache_ok(self, positives_i, negatives_i):
        if self.HOST_CACHE or not torch.is_tensor(positives_i) or not positives_i.is_cuda:
            return False
        return all(torch.is_tensor(b) and b.is_cuda and b.dtype == torch.float32 and b.dim() == 2 for b in negatives_i) \
            and positives_i.dtype == torch.float32

    class _GpuCache:
        def __init__(self, pos, neg_batches):
            from odf import ops
            self.ops = ops
            self.P, self.d = int(pos.shape[0]), int(pos.shape[1])
            cap = self.P + sum(int(b.shape[0]) for b in neg_batches)
            self.buf = [torch.empty((cap, self.d), dtype=torch.float32, device=pos.device) for _ in range(2)]
            self.cur = 0
            self.buf[0][:self.P].copy_(pos)
            self.buf[1][:self.P].copy_(pos)
            self.n = 0

        @property
        def X(self):
            return self.buf[self.cur][:self.P + self.n]

        @property
        def neg(self):
            return self.buf[self.cur][self.P:self.P + self.n]

        def append_all(self, batch):
            k = int(batch.shape[0])
            self.buf[self.cur][self.P + self.n:self.P + self.n + k].copy_(batch)
            self.n += k

        def append_selected(self, batch, scores, thresh):
            idx, cnt = self.ops.select_indices(scores, thresh, strict=True)          # scores > HARD_THRESH
            self.ops.gather_rows(batch, idx, cnt, self.buf[self.cur][self.P + self.n:], max_rows=int(batch.shape[0]))
            k = int(cnt.item())
            self.n += k
            return k, idx[:k]

        def keep_selected(self, scores, thresh):
            idx, cnt = self.ops.select_indices(scores, thresh, strict=False)         # scores >= EASY_THRESH
            other = 1 - self.cur
            self.ops.gather_rows(self.neg, idx, cnt, self.buf[other][self.P:], max_rows=self.n)
            k = int(cnt.item())
            removed = self.n - k
            self.cur, self.n = other, k
            return removed, idx[:k]

    def _train_class_gpu_cache(self, i, positives_i, negatives_i):
        """One class of trainWithMinibootstrap on the pre-allocated GPU cache (same decisions, same order of
        rows, same prints as the generic path below)."""
        cache = self._GpuCache(positives_i, negatives_i)
        model = None
        n_batches = len(negatives_i)
        ones = torch.ones(cache.P, device=positives_i.device)
        for j in range(n_batches):
            t_iter = time.time()
            last = j == n_batches - 1
            if j == 0:
                cache.append_all(negatives_i[0])
            else:
                t_hard = time.time()
                scores = self.classifier.predict(model, negatives_i[j])
                k, _ = cache.append_selected(negatives_i[j], scores, self.hard_tresh)
                print("Hard negatives selected in {} seconds".format(time.time() - t_hard))
                print("Chosen {} hard negatives from the {}th batch".format(k, j))
            print("Traning with {} positives and {} negatives".format(cache.P, cache.n))
            t_update = time.time()
            y = torch.cat((ones, -torch.ones(cache.n, device=ones.device)), 0)
            if self.sigma is not None and self.lam is not None:
                print("Updating model with lambda: {} and sigma: {}".format(self.lam, self.sigma))
                model = self.classifier.train(cache.X, y, sigma=self.sigma, lam=self.lam)
            else:
                print("Updating model with default lambda and sigma")
                model = self.classifier.train(cache.X, y)
            print("Model updated in {} seconds".format(time.time() - t_update))
            t_easy = time.time()
            if cache.n != 0 and not last:
                scores = self.classifier.predict(model, cache.neg)
                removed, _ = cache.keep_selected(scores, self.easy_tresh)
                print("Easy negatives selected in {} seconds".format(time.time() - t_easy))
                print("Removed {} easy negatives. {} Remaining".format(removed, cache.n))
                print("Iteration {}th done in {} seconds".format(j, time.time() - t_iter))
        out_cache = {"pos": positives_i, "neg": cache.neg.clone()} if self.return_caches else None
        return model, out_cache

    # ------------------------------------------------------------------ minibootstrap
    def trainWithMinibootstrap(self, negatives, positives, output_dir=None):
        caches, model = [], []
        t_start = time.time()
        for i in range(self.num_classes - 1):
            if len(positives[i]) == 0 or len(negatives[i]) == 0:
                model.append(None)
                caches.append({})
                continue
            print("---------------------- Training Class number {} ----------------------".format(i))
            if self._gpu_cache_ok(positives[i], negatives[i]):
                m, c = self._train_class_gpu_cache(i, positives[i], negatives[i])
                model.append(m)
                caches.append(c)
                continue
            n_batches = len(negatives[i])
            for j in range(n_batches):
                t_iter = time.time()
                last = j == n_batches - 1
                if j == 0:
                    caches.append({"pos": self._stage(positives[i]), "neg": self._stage(negatives[i][0])})
                    model.append(None)
                else:
                    t_hard = time.time()
                    batch = self._stage(negatives[i][j])
                    scores = self.classifier.predict(model[i], batch)
                    hard_idx = torch.where(scores > self.hard_tresh)[0]
                    caches[i]["neg"] = torch.cat((caches[i]["neg"], batch[hard_idx]), 0)
                    print("Hard negatives selected in {} seconds".format(time.time() - t_hard))
                    print("Chosen {} hard negatives from the {}th batch".format(len(hard_idx), j))
                print("Traning with {} positives and {} negatives".format(len(caches[i]["pos"]), len(caches[i]["neg"])))
                t_update = time.time()
                model[i] = self.updateModel(caches[i])
                print("Model updated in {} seconds".format(time.time() - t_update))
                t_easy = time.time()
                prune = len(caches[i]["neg"]) != 0 and (self.HOST_CACHE or not last)
                if prune:
                    scores = self.classifier.predict(model[i], caches[i]["neg"])
                    keep_idx = torch.where(scores >= self.easy_tresh)[0]
                    removed = len(caches[i]["neg"]) - len(keep_idx)
                    caches[i]["neg"] = caches[i]["neg"][keep_idx]
                    print("Easy negatives selected in {} seconds".format(time.time() - t_easy))
                    print("Removed {} easy negatives. {} Remaining".format(removed, len(caches[i]["neg"])))
                    print("Iteration {}th done in {} seconds".format(j, time.time() - t_iter))
                if last and not self.return_caches:
                    caches[i] = None        # free the cache of a finished class
                    if torch.cuda.is_available():
                        torch.cuda.empty_cache()
        training_time = time.time() - t_start
        print("Online Classifier trained in {} seconds".format(training_time))
        if output_dir:
            if self.is_rpn:
                head = "RPN's Online Classifier training time"
            elif self.is_segmentation:
                head = "Online Segmentation training time"
            else:
                head = "Detector's Online Classifier training time"
            with open(os.path.join(output_dir, "result.txt"), "a") as fid:
                fid.write("{}: {}min:{}s \n".format(head, int(training_time / 60), round(training_time % 60)))
        if self.return_caches:
            self.caches = caches
        return model

    def trainRegionClassifier(self, opts=None, output_dir=None):
        if opts is not None:
            self.processOptions(opts)
        print("Training Online Region Classifier")
        negatives, positives = self.negatives, self.positives
        if self.HOST_CACHE:
            dev = negatives[0][0].device
            self.mean, self.std, self.mean_norm = self.mean.to(dev), self.std.to(dev), self.mean_norm.to(dev)
        if not self.normalized:
            for i in range(self.num_classes - 1):
                if len(positives[i]):
                    positives[i] = self.zScores(positives[i])
                for j in range(len(negatives[i])):
                    if len(negatives[i][j]):
                        negatives[i][j] = self.zScores(negatives[i][j])
            self.normalized = True
        model = self.trainWithMinibootstrap(negatives, positives, output_dir=output_dir)
        if self.return_caches:
            return model, self.caches
        return model

    # ------------------------------------------------------------------ per-image scoring
    def _parallel_head(self, model):
        """nystrom_parallel / alpha_parallel of the trained classes, as the reference's inference heads build them
        (roi_box_predictors.py:140-160): the centres of all classes stacked, alpha block-diagonal (one column per class;
        a class without a model keeps an all-zero column and is reported as -1 below, as the sequential loop would leave
        it).  One call of kernel.mmv then scores every class of an image: one operand pre-pass (with the z-score fused)
        and one launch of the fused tile instead of one of each per class."""
        live = [(c, m) for c, m in enumerate(model[:self.num_classes - 1]) if m is not None]
        if not live:
            return None
        dev = torch.device("cuda", torch.cuda.current_device())
        total = sum(int(m.ny_points_.shape[0]) for _c, m in live)
        alpha = torch.zeros((total, self.num_classes - 1), dtype=torch.float32, device=dev)
        row = 0
        for c, m in live:
            k = int(m.ny_points_.shape[0])
            alpha[row:row + k, c] = m.alpha_.to(dev).reshape(-1)
            row += k
        centres = torch.cat([m.ny_points_.to(device=dev, dtype=torch.float32) for _c, m in live])
        return live[0][1].kernel, centres, alpha, [c for c, _m in live]

    def testRegionClassifier(self, model, test_boxes):
        print("Online Region Classifier testing")
        predictions = []
        total = 0.0
        try:
            for c in range(self.num_classes - 1):
                model[c].ny_points_ = model[c].ny_points_.to("cuda")
                model[c].alpha_ = model[c].alpha_.to("cuda")
        except Exception:  # noqa: BLE001  (None models: reference swallows this too)
            pass
        if self.HOST_CACHE and torch.is_tensor(self.mean):
            self.mean, self.std, self.mean_norm = self.mean.to("cuda"), self.std.to("cuda"), self.mean_norm.to("cuda")
        # libodf models (their kernel caches the stacked operands between images): score all classes in one call
        odf_models = any(m is not None and hasattr(getattr(m, "kernel", None), "_cached") for m in model)
        head = self._parallel_head(model) if odf_models else None
        for entry in test_boxes:
            if entry is None:
                continue
            not_gt = np.nonzero(entry["gt"] == 0)
            boxes = entry["boxes"][not_gt, :][0]
            X_test = torch.tensor(entry["feat"][not_gt, :][0], device="cuda")
            t0 = time.time()
            scores = -torch.ones((len(boxes), self.num_classes))
            if head is not None:
                kernel, centres, alpha, live = head
                zs = None
                if self.mean_norm != 0:
                    zs = (self.mean.to(device=X_test.device, dtype=torch.float32), 20.0 / self.mean_norm.item())
                s = kernel.mmv(X_test.float(), centres, alpha, zscore=zs)
                scores[:, [c + 1 for c in live]] = s[:, live].cpu()
            else:
                if self.mean_norm != 0:
                    X_test = self.zScores(X_test)
                for c in range(self.num_classes - 1):
                    scores[:, c + 1] = torch.squeeze(self.classifier.predict(model[c], X_test))
            total += time.time() - t0
            b = BoxList(torch.from_numpy(boxes), (entry["img_size"][0], entry["img_size"][1]), mode="xyxy")
            b.add_field("scores", scores.to("cpu"))
            predictions.append(b)
        print("Average image testing time: {} seconds.".format(total / max(len(test_boxes), 1)))
        return predictions

    def predict(self, dataset):
        pass

    def zScores(self, feat, target_norm=20):
        feat = feat - self.mean
        return feat * (target_norm / self.mean_norm.item())
