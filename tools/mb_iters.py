"""How many CG iterations do the refits of the minibootstrap workload run (does the early exit on convergence matter)?"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

import torch  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "online-detection_b200"))
import odf  # noqa: E402

counts = collections.Counter()
orig = odf.InCoreFalkon.fit


def fit(self, *a, **k):
    r = orig(self, *a, **k)
    counts[self.fit_times_["cg_iters"]] += 1
    return r


odf.InCoreFalkon.fit = fit
sys.argv = ["bench.py", "--workload", "mb", "--n", "3", "--steps", "1", "--warmup", "1"]
bench.main()
print("cg iterations per refit:", dict(sorted(counts.items())), file=sys.stderr)
