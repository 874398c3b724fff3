"""Nodes from latitude / longitude arrays (/root/reference/src/anemoi/graphs/nodes/builders/from_vectors.py)."""

from __future__ import annotations

import numpy as np
import torch

from .base import BaseNodeBuilder


class LatLonNodes(BaseNodeBuilder):
    """Nodes from its latitude and longitude positions (in numpy arrays), in degrees."""

    def __init__(self, latitudes: list[float] | np.ndarray, longitudes: list[float] | np.ndarray, name: str) -> None:
        super().__init__(name)
        self.latitudes = latitudes if isinstance(latitudes, np.ndarray) else np.array(latitudes)
        self.longitudes = longitudes if isinstance(longitudes, np.ndarray) else np.array(longitudes)

        assert len(self.latitudes) == len(
            self.longitudes
        ), f"Lenght of latitudes and longitudes must match but {len(self.latitudes)}!={len(self.longitudes)}."
        assert self.latitudes.ndim == 1 or (
            self.latitudes.ndim == 2 and self.latitudes.shape[1] == 1
        ), "latitudes must have shape (N, ) or (N, 1)."
        assert self.longitudes.ndim == 1 or (
            self.longitudes.ndim == 2 and self.longitudes.shape[1] == 1
        ), "longitudes must have shape (N, ) or (N, 1)."

    def get_coordinates(self) -> torch.Tensor:
        return self.reshape_coords(self.latitudes, self.longitudes)
