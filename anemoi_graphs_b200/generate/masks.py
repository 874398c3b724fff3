"""Area-of-interest mask from the distance to a reference node set
(/root/reference/src/anemoi/graphs/generate/masks.py:23-99).

The reference fits a haversine ``NearestNeighbors`` and tests ``d_NN * 6371 <= margin_radius_km``; here the
k=1 query runs on the GPU index and the float64 distance ``2 asin(sqrt(rdist))`` is evaluated on the
device."""

from __future__ import annotations

import logging

import numpy as np
import torch

from .. import EARTH_RADIUS
from .. import device as _device
from .. import ops

LOGGER = logging.getLogger(__name__)


class KNNAreaMaskBuilder:
    """Class to build a mask based on distance to masked reference nodes using KNN.

    Attributes
    ----------
    margin_radius_km : float
        Maximum distance to the reference nodes to consider a node as valid, in kilometers. Defaults to 100 km.
    reference_node_name : str
        Name of the reference nodes in the graph to consider for the Area Mask.
    mask_attr_name : str
        Name of a node to attribute to mask the reference nodes, if desired. Defaults to consider all reference nodes.
    """

    def __init__(self, reference_node_name: str, margin_radius_km: float = 100, mask_attr_name: str | None = None):
        assert isinstance(margin_radius_km, (int, float)), "The margin radius must be a number."
        assert margin_radius_km > 0, "The margin radius must be positive."

        self.margin_radius_km = margin_radius_km
        self.reference_node_name = reference_node_name
        self.mask_attr_name = mask_attr_name
        self._reference: torch.Tensor | None = None  # CUDA float32 (n, 2)

    def get_reference_coords(self, graph) -> torch.Tensor:
        """Retrieve (device) coordinates of the reference nodes."""
        assert (
            self.reference_node_name in graph.node_types
        ), f'Reference node "{self.reference_node_name}" not found in the graph.'

        nodes = graph[self.reference_node_name]
        coords_rad = _device.node_state(nodes).x
        if self.mask_attr_name is not None:
            assert (
                self.mask_attr_name in nodes.node_attrs()
            ), f'Mask attribute "{self.mask_attr_name}" not found in the reference nodes.'
            mask = nodes[self.mask_attr_name].squeeze().to(device=coords_rad.device, dtype=torch.bool)
            coords_rad = coords_rad[mask]

        return coords_rad

    def fit_coords(self, coords_rad) -> None:
        """Fit to the coordinates in radians (numpy array, CPU or CUDA tensor)."""
        if isinstance(coords_rad, np.ndarray):
            coords_rad = torch.from_numpy(np.ascontiguousarray(coords_rad))
        self._reference = _device.to_device(coords_rad, torch.float32)
        self.n_samples_fit_ = int(self._reference.shape[0])

    def fit(self, graph) -> None:
        """Fit to the nodes of interest."""
        reference_mask_str = self.reference_node_name
        if self.mask_attr_name is not None:
            reference_mask_str += f" ({self.mask_attr_name})"

        coords_rad = self.get_reference_coords(graph)
        self.fit_coords(coords_rad)

        LOGGER.info(
            'Fitting %s with %d reference nodes from "%s".',
            self.__class__.__name__,
            len(coords_rad),
            reference_mask_str,
        )

    def get_mask_device(self, coords_rad: torch.Tensor) -> torch.Tensor:
        """CUDA bool mask: nearest reference node within ``margin_radius_km``."""
        assert self._reference is not None, f"{self.__class__.__name__} must be fitted first."
        q = _device.to_device(coords_rad, torch.float32)
        # only "is the nearest reference node within the margin?" is asked, so the search stops a little beyond it -
        # a query on the far side of the globe from a limited-area patch must not walk every cell on the way
        limit = 1.01 * self.margin_radius_km / EARTH_RADIUS + 1e-6
        with ops.NeighbourIndex(self._reference, hint_k=1) as index:
            _, rdist = index.knn(q, 1, return_rdist=True, tag="knn_mask", max_radius=limit)
        dist = 2.0 * torch.asin(torch.sqrt(rdist[:, 0]))  # HaversineDistance64._rdist_to_dist
        return dist * EARTH_RADIUS <= self.margin_radius_km

    def get_mask(self, coords_rad) -> np.ndarray:
        """Compute a mask based on the distance to the reference nodes (generate/masks.py:94-99)."""
        if isinstance(coords_rad, np.ndarray):
            coords_rad = torch.from_numpy(np.ascontiguousarray(coords_rad))
        return self.get_mask_device(coords_rad).cpu().numpy()
