"""Drop-in for src/modules/accuracy-evaluator/OnlineDetectionPostProcessor_standalone.py: same class,
`forward(boxes, num_classes)` takes already-decoded per-image BoxLists (reference :11-60)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.dirname(__file__)))
from OnlineDetectionPostProcessor import OnlineDetectionPostProcessor as _Base  # noqa: E402


class OnlineDetectionPostProcessor(_Base):
    def forward(self, boxes, num_classes):
        return self.forward_standalone(boxes, num_classes)
