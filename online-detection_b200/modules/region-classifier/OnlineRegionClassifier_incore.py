"""Drop-in for src/modules/region-classifier/OnlineRegionClassifier_incore.py (GPU-resident
caches; no easy-negative pruning after the last batch, reference :130)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _region_classifier import BoxList, OnlineRegionClassifierBase  # noqa: E402,F401


class OnlineRegionClassifier(OnlineRegionClassifierBase):
    HOST_CACHE = False
