"""Graph post-processors (/root/reference/src/anemoi/graphs/processors/post_process.py:22-149).

``RemoveUnconnectedNodes`` keeps the reference's constructor, ``compute_mask`` / ``update_graph`` contract and
results; the per-edge work - marking connected nodes and rewriting every edge endpoint through the old -> new
index map, which the reference does with a python dict and ``Tensor.apply_`` - runs on the GPU
(``agx_mark_nodes`` -> ``agx_exclusive_scan`` -> ``agx_relabel_nodes``).  Tensors come back where they were
(CPU in, CPU out), with their dtypes.
"""

from __future__ import annotations

import logging
from abc import ABC
from abc import abstractmethod

import torch

from .. import device as _device
from .._cabi import check, current_stream, load_library, ptr

LOGGER = logging.getLogger(__name__)


class PostProcessor(ABC):
    _agx_device_aware = True

    @abstractmethod
    def update_graph(self, graph):
        raise NotImplementedError(f"The {self.__class__.__name__} class does not implement the method update_graph().")


def _row_int32(row: torch.Tensor) -> torch.Tensor:
    """One edge_index row as a contiguous CUDA int32 tensor (the row itself when it already is one)."""
    return _device.to_device(row.contiguous(), torch.int32)


class BaseMaskingProcessor(PostProcessor, ABC):
    """Base class for mask based processor."""

    def __init__(self, nodes_name: str, save_mask_indices_to_attr: str | None = None) -> None:
        self.nodes_name = nodes_name
        self.save_mask_indices_to_attr = save_mask_indices_to_attr
        self.mask: torch.Tensor = None
        self._flags_dev: torch.Tensor | None = None  # CUDA int32 keep flags behind ``mask``

    def removing_nodes(self, graph):
        """Remove nodes based on the mask passed."""
        nodes = graph[self.nodes_name]
        for attr_name in nodes.node_attrs():
            value = nodes[attr_name]
            nodes[attr_name] = value[self.mask.to(value.device)]
        return graph

    def create_indices_mapper_from_mask(self) -> dict[int, int]:
        return dict(zip(torch.where(self.mask)[0].tolist(), list(range(int(self.mask.sum())))))

    def update_edge_indices(self, graph):
        """Update the edge indices to the new position of the nodes (post_process.py:49-60)."""
        lib = load_library()
        flags = self._flags_dev
        if flags is None:
            flags = _device.to_device(self.mask, torch.int32)
        n_nodes = int(flags.shape[0])
        new_index = torch.empty(n_nodes + 1, dtype=torch.int64, device=flags.device)
        stream = current_stream()
        check(lib.agx_exclusive_scan(ptr(flags), n_nodes, ptr(new_index), None, stream))
        for edges_name in graph.edge_types:
            for row, name in ((0, edges_name[0]), (1, edges_name[2])):
                if name != self.nodes_name:
                    continue
                edge_index = graph[edges_name].edge_index
                dev_row = _row_int32(edge_index[row])
                check(lib.agx_relabel_nodes(ptr(dev_row), int(dev_row.shape[0]), ptr(new_index), stream))
                edge_index[row] = dev_row.to(device=edge_index.device, dtype=edge_index.dtype)
        return graph

    @abstractmethod
    def compute_mask(self, graph) -> torch.Tensor: ...

    def add_attribute(self, graph):
        """Add an attribute of the mask indices as node attribute."""
        if self.save_mask_indices_to_attr is not None:
            LOGGER.info(
                f"An attribute {self.save_mask_indices_to_attr} has been added with the indices to mask the nodes from the original graph."
            )
            mask_indices = torch.where(self.mask)[0].reshape((graph[self.nodes_name].num_nodes, -1))
            graph[self.nodes_name][self.save_mask_indices_to_attr] = mask_indices
        return graph

    def update_graph(self, graph):
        """Post-process the graph."""
        self.mask = self.compute_mask(graph)
        LOGGER.info(f"Removing {(~self.mask).sum()} nodes from {self.nodes_name}.")
        graph = self.removing_nodes(graph)
        graph = self.update_edge_indices(graph)
        graph = self.add_attribute(graph)
        self._flags_dev = None
        return graph


class RemoveUnconnectedNodes(BaseMaskingProcessor):
    """Remove unconnected nodes in the graph.

    Attributes
    ----------
    nodes_name: str
        Name of the unconnected nodes to remove.
    ignore: str, optional
        Name of an attribute to ignore when removing nodes. Nodes with
        this attribute set to True will not be removed.
    save_mask_indices_to_attr: str, optional
        Name of the attribute to save the mask indices. If provided,
        the indices of the kept nodes will be saved in this attribute.
    """

    def __init__(
        self,
        nodes_name: str,
        save_mask_indices_to_attr: str | None = None,
        ignore: str | None = None,
    ) -> None:
        super().__init__(nodes_name, save_mask_indices_to_attr)
        self.ignore = ignore

    def compute_mask(self, graph) -> torch.Tensor:
        """Compute the mask of connected nodes (post_process.py:133-149): bool (num_nodes,), on the device ``x``
        lives on."""
        nodes = graph[self.nodes_name]
        n_nodes = int(nodes.num_nodes)
        dev = _device.compute_device()
        flags = torch.zeros(n_nodes, dtype=torch.int32, device=dev)

        if self.ignore is not None:
            LOGGER.info(f"The nodes with {self.ignore}=True will not be removed.")
            flags[nodes[self.ignore].bool().squeeze().to(dev)] = 1

        lib = load_library()
        stream = current_stream()
        for (source_name, _, target_name), edges in graph.edge_items():
            for row, name in ((0, source_name), (1, target_name)):
                if name != self.nodes_name:
                    continue
                dev_row = _row_int32(edges.edge_index[row])
                check(lib.agx_mark_nodes(ptr(dev_row), int(dev_row.shape[0]), n_nodes, ptr(flags), stream))

        self._flags_dev = flags
        return flags.bool().to(nodes["x"].device)
