from .builder import CutOffEdges
from .builder import KNNEdges
from .builder import MultiScaleEdges

__all__ = ["KNNEdges", "CutOffEdges", "MultiScaleEdges"]
