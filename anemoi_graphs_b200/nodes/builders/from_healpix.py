"""Nodes from the HEALPix grid (/root/reference/src/anemoi/graphs/nodes/builders/from_healpix.py:23-111).

The reference asks healpy for the nested pixel centres; here one kernel evaluates HEALPix's pix2loc
(``ops.healpix_nodes``)."""

from __future__ import annotations

import logging
import math

import torch

from ... import device as _device
from ... import ops
from ...generate.masks import KNNAreaMaskBuilder
from .base import BaseNodeBuilder

LOGGER = logging.getLogger(__name__)


class HEALPixNodes(BaseNodeBuilder):
    """Nodes from HEALPix grid (Hierarchical Equal Area isoLatitude Pixelization of a sphere).

    Attributes
    ----------
    resolution : int
        The resolution of the grid (nside = 2**resolution).
    """

    def __init__(self, resolution: int, name: str) -> None:
        self.resolution = resolution
        super().__init__(name)

        assert isinstance(resolution, int), "Resolution must be an integer."
        assert resolution > 0, "Resolution must be positive."

    def _pixel_centres(self) -> torch.Tensor:
        nside = 2**self.resolution
        spatial_res_degrees = math.degrees(math.sqrt(4.0 * math.pi / (12 * nside * nside)))  # hp.nside2resol
        LOGGER.info(f"Creating HEALPix nodes with resolution {spatial_res_degrees:.2} deg.")
        return ops.healpix_nodes(self.resolution)

    def get_coordinates(self) -> torch.Tensor:
        """float32 (num_nodes, 2) coordinates of the nodes, in radians."""
        self._x_device = self._pixel_centres()
        return self._x_device if _device.is_resident() else _device.to_host(self._x_device)

    def register_nodes(self, graph):
        graph = super().register_nodes(graph)
        _device.seed_node_state(graph[self.name], self._x_device)
        _device.maybe_flush()
        return graph


class LimitedAreaHEALPixNodes(HEALPixNodes):
    """Nodes from HEALPix grid using an area of interest.

    In the reference this class cannot be used: its constructor assigns ``area_mask_builder`` before calling the
    base constructor, which resets it to None (from_healpix.py:84-87, nodes/builders/base.py:38), so
    ``register_nodes`` raises AttributeError.  The evident intent (from_healpix.py:93-110) is implemented here."""

    def __init__(
        self,
        resolution: int,
        reference_node_name: str,
        name: str,
        mask_attr_name: str | None = None,
        margin_radius_km: float = 100.0,
    ) -> None:
        super().__init__(resolution, name)
        self.area_mask_builder = KNNAreaMaskBuilder(reference_node_name, margin_radius_km, mask_attr_name)

    def register_nodes(self, graph):
        self.area_mask_builder.fit(graph)
        return super().register_nodes(graph)

    def get_coordinates(self) -> torch.Tensor:
        coords = self._pixel_centres()
        LOGGER.info(
            'Limiting the "%s" nodes to a radius of %.2f km from the nodes of interest.',
            self.name,
            self.area_mask_builder.margin_radius_km,
        )
        area_mask = self.area_mask_builder.get_mask_device(coords)
        LOGGER.info('Masking out %d nodes from "%s".', int(area_mask.numel() - area_mask.sum().item()), self.name)
        self._x_device = coords[area_mask]
        return self._x_device if _device.is_resident() else _device.to_host(self._x_device)
