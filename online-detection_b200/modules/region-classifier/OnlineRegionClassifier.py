"""Drop-in for src/modules/region-classifier/OnlineRegionClassifier.py ("--CPU" flavour: caches
parked in host RAM, easy negatives pruned after every refit, reference :108-137)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _region_classifier import BoxList, OnlineRegionClassifierBase  # noqa: E402,F401


class OnlineRegionClassifier(OnlineRegionClassifierBase):
    HOST_CACHE = True
