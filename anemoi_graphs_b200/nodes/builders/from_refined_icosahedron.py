"""Nodes from iterative refinements of an icosahedron
(/root/reference/src/anemoi/graphs/nodes/builders/from_refined_icosahedron.py:30-191).

Same class names, constructor arguments and hidden attributes (``_resolutions``, ``_nx_graph``,
``_node_ordering``, ``_area_mask_builder``) as the reference; ``_nx_graph`` holds the device mesh
(``generate.tri_icosahedron.DeviceMesh``) instead of a networkx graph.
"""

from __future__ import annotations

import logging
from abc import ABC
from abc import abstractmethod

import numpy as np
import torch

from ... import device as _device
from ...generate.hex_icosahedron import create_hex_nodes
from ...generate.masks import KNNAreaMaskBuilder
from ...generate.tri_icosahedron import create_stretched_tri_nodes
from ...generate.tri_icosahedron import create_tri_nodes
from ...generate.tri_icosahedron import create_tri_nodes_provisional
from .base import BaseNodeBuilder

LOGGER = logging.getLogger(__name__)



class IcosahedralNodes(BaseNodeBuilder, ABC):
    """Nodes based on iterative refinements of an icosahedron.

    Attributes
    ----------
    resolution : list[int] | int
        Refinement level of the mesh.
    """

    def __init__(self, resolution: int | list[int], name: str) -> None:
        if isinstance(resolution, int):
            self.resolutions = list(range(resolution + 1))
        else:
            self.resolutions = resolution

        super().__init__(name)
        self.hidden_attributes = BaseNodeBuilder.hidden_attributes | {
            "resolutions",
            "nx_graph",
            "node_ordering",
            "area_mask_builder",
        }

    def get_coordinates(self) -> torch.Tensor:
        """float32 (num_nodes, 2) coordinates in radians, in graph order."""
        self._provisional = None
        if self.provisional_order and _device.deferring() and _device.LAZY_NODE_ORDER:
            # inside a deferred scope (GraphCreator.update_graph): the order is sorted on a host thread while the
            # edge kernels already run in the generator's numbering; x / _node_ordering are complete at flush
            self.nx_graph, prov = self.create_nodes_provisional()
            self._provisional = prov
            self._x_device = prov.x_final
            self.node_ordering = prov.order_host.numpy()
            if _device.is_resident():
                return prov.x_final
            prov.x_host = _device.host_tensor((prov.n, 2), torch.float32)
            return prov.x_host
        self.nx_graph, coords_rad, order = self.create_nodes()
        # == torch.tensor(coords_rad[node_ordering], dtype=torch.float32), gathered (and, for the float64 hexagonal
        # centres, rounded) on the device
        self._x_device = coords_rad[order].to(torch.float32)
        # the hidden ``_node_ordering`` attribute is a host array in the reference: a pinned copy, complete when the
        # enclosing deferred scope (or this builder's register_nodes) flushes
        self.node_ordering = _device.to_host(order).numpy()
        return self._x_device if _device.is_resident() else _device.to_host(self._x_device)

    def register_nodes(self, graph):
        graph = super().register_nodes(graph)
        nodes = graph[self.name]
        # the device copy already exists: seed the per-node-set state so no builder uploads x again
        st = _device.seed_node_state(nodes, self._x_device)
        if self._provisional is not None:
            self._provisional.attach(nodes, st)
        _device.maybe_flush()
        return graph

    provisional_order = False  # subclasses whose order is host-sorted and that can defer it set this

    def create_nodes_provisional(self):
        raise NotImplementedError

    @abstractmethod
    def create_nodes(self) -> tuple[object, torch.Tensor, torch.Tensor]: ...


class LimitedAreaIcosahedralNodes(IcosahedralNodes):
    """Icosahedral nodes restricted to an area of interest."""

    def __init__(
        self,
        resolution: int | list[int],
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


class TriNodes(IcosahedralNodes):
    """Nodes based on iterative refinements of an icosahedron (triangular mesh)."""

    provisional_order = True

    def create_nodes(self):
        return create_tri_nodes(resolution=max(self.resolutions))

    def create_nodes_provisional(self):
        return create_tri_nodes_provisional(resolution=max(self.resolutions))


class HexNodes(IcosahedralNodes):
    """Nodes based on iterative refinements of an icosahedron (H3 hexagonal cells).

    The reference depends on the h3 Python library; here the cells come from ``agx_hex_cells``."""

    def create_nodes(self):
        return create_hex_nodes(resolution=max(self.resolutions))


class LimitedAreaTriNodes(LimitedAreaIcosahedralNodes):
    """Triangular-mesh nodes within an area of interest."""

    def create_nodes(self):
        return create_tri_nodes(resolution=max(self.resolutions), area_mask_builder=self.area_mask_builder)


class LimitedAreaHexNodes(LimitedAreaIcosahedralNodes):
    """H3 hexagonal-cell nodes within an area of interest."""

    def create_nodes(self):
        return create_hex_nodes(resolution=max(self.resolutions), area_mask_builder=self.area_mask_builder)


class StretchedIcosahedronNodes(IcosahedralNodes):
    """Icosahedral nodes with 2 different resolutions."""

    def __init__(
        self,
        global_resolution: int,
        lam_resolution: int,
        name: str,
        reference_node_name: str,
        mask_attr_name: str,
        margin_radius_km: float = 100.0,
    ) -> None:
        super().__init__(lam_resolution, name)
        self.global_resolution = global_resolution

        self.area_mask_builder = KNNAreaMaskBuilder(reference_node_name, margin_radius_km, mask_attr_name)

    def register_nodes(self, graph):
        self.area_mask_builder.fit(graph)
        return super().register_nodes(graph)


class StretchedTriNodes(StretchedIcosahedronNodes):
    """Triangular-mesh nodes with 2 different resolutions."""

    def create_nodes(self):
        return create_stretched_tri_nodes(
            base_resolution=self.global_resolution,
            lam_resolution=max(self.resolutions),
            area_mask_builder=self.area_mask_builder,
        )
