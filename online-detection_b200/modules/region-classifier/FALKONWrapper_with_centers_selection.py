"""Drop-in for src/modules/region-classifier/FALKONWrapper_with_centers_selection.py
(host-RAM "--CPU" flavour, reference :16-95): features may live on the host; the fit and the
scoring still run on the GPU, results come back on the caller's device."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _falkon_wrapper import FALKONWrapperBase, Falkon, kernels  # noqa: E402,F401
from _falkon_wrapper import FalkonOptions, MyCenterSelector  # noqa: E402,F401


class FALKONWrapper(FALKONWrapperBase):
    MODEL_CLS = Falkon
    IN_CORE = False
