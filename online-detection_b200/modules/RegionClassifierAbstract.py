"""Abstract region-classifier surface (reference: src/modules/region-classifier/
RegionClassifierAbstract.py:9-41)."""
import abc


class RegionClassifierAbstract(abc.ABC):
    @abc.abstractmethod
    def loadRegionClassifier(self):
        ...

    @abc.abstractmethod
    def trainRegionClassifier(self, dataset):
        ...

    @abc.abstractmethod
    def testRegionClassifier(self, dataset):
        ...

    @abc.abstractmethod
    def predict(self, dataset):
        ...
