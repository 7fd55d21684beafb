"""Nystrom-centre selector handed to the FALKON model
(reference: src/modules/region-classifier/MyCenterSelector.py:3-15)."""
import torch


class MyCenterSelector:
    """Returns the rows of X (and Y) at a fixed list of indices; duplicates are allowed."""

    def __init__(self, center_indices):
        self.center_indices = center_indices

    def _index(self, ref):
        idx = self.center_indices
        if not torch.is_tensor(idx):
            idx = torch.as_tensor(idx, dtype=torch.int64)
        return idx.to(ref.device)

    def select(self, X, Y):
        idx = self._index(X)
        centres = X.index_select(0, idx.reshape(-1))
        if centres.dim() > 2:
            centres = centres.squeeze()
        if Y is None:
            return centres
        return centres, Y.index_select(0, idx.reshape(-1))
