"""Make the `odf` host package importable from the drop-in module directories (the reference's
scripts reach these files through sys.path.append, experiments/run_experiment_*.py:6-12)."""
import os
import sys

_PKG = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
_MOD = os.path.abspath(os.path.dirname(__file__))
if _MOD not in sys.path:
    sys.path.insert(0, _MOD)
