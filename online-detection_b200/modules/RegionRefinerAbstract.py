"""Abstract region-refiner surface (reference: src/modules/RegionRefinerAbstract.py)."""
import abc


class RegionRefinerAbstract(abc.ABC):
    @abc.abstractmethod
    def loadRegionRefiner(self):
        ...

    @abc.abstractmethod
    def trainRegionRefiner(self):
        ...

    @abc.abstractmethod
    def testRegionRefiner(self):
        ...

    @abc.abstractmethod
    def predict(self):
        ...
