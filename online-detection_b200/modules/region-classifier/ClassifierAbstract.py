"""Abstract classifier surface (reference: src/modules/region-classifier/ClassifierAbstract.py:4-18)."""
import abc


class ClassifierAbstract(abc.ABC):
    """train / predict / test — the three verbs every on-line classifier wrapper provides."""

    @abc.abstractmethod
    def train(self, dataset):
        ...

    @abc.abstractmethod
    def predict(self, dataset):
        ...

    @abc.abstractmethod
    def test(self, dataset):
        ...
