"""Golden vectors for the data formats either side of the hot path: the REFERENCE's own loaders and helpers
(src/py_od_utils.py: load_features_classifier :120-200 incl. the is_segm / cpu_tensor / sample_ratio / shuffled variants,
load_features_regressor :202-224, minibatch_positives :241-245, mask_iou :297-331, zScores :98-103) run on the
deterministic feature caches of tests/golden/format_fixture.py.  Output: tests/golden/reference_formats.npz, checked
against the product's drop-in py_od_utils.py by tests/test_reference_golden.py (CPU).

    python tests/golden/make_reference_golden_formats.py      # needs /root/reference (not present on the GPU box)
"""
import contextlib
import io
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import format_fixture as fx  # noqa: E402
import make_reference_golden as base  # noqa: E402


def main():
    if not os.path.isdir(os.path.join(base.REF, "src")):
        raise SystemExit("reference tree not found at %s" % base.REF)
    sys.path.insert(0, os.path.join(base.REF, "src"))
    with tempfile.TemporaryDirectory() as root, base.cuda_is_cpu(), contextlib.redirect_stdout(io.StringIO()), \
            warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfg = fx.build(root)
        import py_od_utils as UT                                   # the reference's file
        assert UT.__file__.startswith(base.REF)
        out = fx.run_all(UT, root, cfg)
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, "reference_formats.npz"), **arrays)
    print("wrote reference_formats.npz (%d arrays)" % len(arrays))
    print("det classes", int(arrays["det_n_classes"][0]), "pos1 rows", arrays["det_pos1"].shape, "shuffled batches",
          int(arrays["det_shuffled_neg0_n"][0]), "mask_iou", arrays["mask_iou"].shape)


if __name__ == "__main__":
    main()
