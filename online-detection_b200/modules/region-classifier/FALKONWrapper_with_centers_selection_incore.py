"""Drop-in for src/modules/region-classifier/FALKONWrapper_with_centers_selection_incore.py
(GPU-resident flavour, reference :16-99).  `import FALKONWrapper_with_centers_selection_incore
as falkon; falkon.FALKONWrapper(cfg_path, is_rpn=...)` works unchanged."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _falkon_wrapper import FALKONWrapperBase, InCoreFalkon, kernels  # noqa: E402,F401
from _falkon_wrapper import FalkonOptions, MyCenterSelector  # noqa: E402,F401


class FALKONWrapper(FALKONWrapperBase):
    MODEL_CLS = InCoreFalkon
    IN_CORE = True
