"""Drop-in for src/modules/region-refiner/region_refiner.py:8-36 (RegionRefiner facade)."""
import os
import sys

import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
sys.path.insert(0, os.path.abspath(os.path.join(_HERE, os.pardir)))
from RegionRefinerAbstract import RegionRefinerAbstract  # noqa: E402
from region_predictor import RegionPredictor  # noqa: E402
from region_refiner_trainer import RegionRefinerTrainer  # noqa: E402


class RegionRefiner(RegionRefinerAbstract):
    def __init__(self, cfg_path_region_refiner, is_rpn=False):
        with open(cfg_path_region_refiner) as fh:
            self.cfg = yaml.load(fh, Loader=yaml.FullLoader)
        if is_rpn:
            self.cfg = self.cfg["RPN"]
        try:
            self.lambd = self.cfg["REGION_REFINER"]["opts"]["lambda"]
        except Exception:  # noqa: BLE001
            self.lambd = None
        self.is_rpn = is_rpn
        self.models = None

    def loadRegionRefiner(self):
        return

    def trainRegionRefiner(self, COXY, output_dir=None):
        trainer = RegionRefinerTrainer(self.cfg, lmbd=self.cfg["REGION_REFINER"]["opts"]["lambda"], is_rpn=self.is_rpn)
        self.models = trainer(COXY, output_dir=output_dir)
        return self.models

    def testRegionRefiner(self):
        return

    def predict(self, boxes, features, models=None, normalize_features=False, stats=None):
        predictor = RegionPredictor(self.cfg, self.models if models is None else models)
        return predictor(boxes, features, normalize_features=normalize_features, stats=stats)
