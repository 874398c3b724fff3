"""Base node builder (/root/reference/src/anemoi/graphs/nodes/builders/base.py:22-125)."""

from __future__ import annotations

from abc import ABC
from abc import abstractmethod

import numpy as np
import torch

from ... import device as _device
from ...config import DotDict
from ...config import instantiate


class BaseNodeBuilder(ABC):
    """Base class for node builders.

    The node coordinates are stored in the `x` attribute of the nodes and they are stored in radians.

    Attributes
    ----------
    name : str
        name of the nodes, key for the nodes in the HeteroData graph object.
    area_mask_builder : KNNAreaMaskBuilder
        The area of interest mask builder, if any. Defaults to None.
    """

    hidden_attributes: set[str] = set()
    _agx_device_aware = True

    def __init__(self, name: str) -> None:
        self.name = name
        self.area_mask_builder = None

    def register_nodes(self, graph):
        """Register nodes in the graph."""
        x = self.get_coordinates()
        if _device.is_resident():
            x = _device.to_device(x, torch.float32)
        graph[self.name].x = x
        graph[self.name].node_type = type(self).__name__
        return graph

    def register_attributes(self, graph, config: DotDict | None = None):
        """Register attributes in the nodes of the graph specified."""
        for hidden_attr in self.hidden_attributes:
            graph[self.name][f"_{hidden_attr}"] = getattr(self, hidden_attr)

        for attr_name, attr_config in (config or {}).items():
            attribute = instantiate(attr_config)
            _device.flush_for(attribute)  # a foreign attribute object may read the host tensors
            graph[self.name][attr_name] = attribute.compute(graph, self.name)

        return graph

    @abstractmethod
    def get_coordinates(self) -> torch.Tensor: ...

    def reshape_coords(self, latitudes: np.ndarray, longitudes: np.ndarray) -> torch.Tensor:
        """Latitude / longitude in degrees, shape (num_nodes,) -> float32 (num_nodes, 2) in radians."""
        coords = np.stack([latitudes, longitudes], axis=-1).reshape((-1, 2))
        coords = np.deg2rad(coords)
        return torch.tensor(coords, dtype=torch.float32)

    def update_graph(self, graph, attrs_config: DotDict | None = None):
        """Update the graph with new nodes."""
        graph = self.register_nodes(graph)

        if attrs_config is None:
            return graph

        graph = self.register_attributes(graph, attrs_config)

        return graph
