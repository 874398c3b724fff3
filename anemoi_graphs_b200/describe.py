"""Graph description (/root/reference/src/anemoi/graphs/describe.py:20-225).

Same ``GraphDescriptor`` surface (``total_size``, ``get_node_summary``, ``get_edge_summary``,
``get_attribute_table``, ``describe``).  The reference reduces every attribute on the host one ``.item()`` at a time
and counts isolated nodes with ``torch.unique`` (a sort of every edge row: 20 M entries for the O1280 decoder); here
the tensors are moved to the GPU once when one is present and the reductions run there.  No custom kernels: this is
bookkeeping around the path, kept so that ``anemoi-graphs describe graph.pt`` has an equivalent.
"""

from __future__ import annotations

import math
from itertools import chain
from pathlib import Path
from typing import Optional
from typing import Union

import torch


def _bytes(n: float) -> str:
    """``anemoi.utils.humanize.bytes``: 1024-based, one decimal."""
    for unit in ("", "KiB", "MiB", "GiB", "TiB"):
        if abs(n) < 1024 or unit == "TiB":
            return f"{n:.0f}" if unit == "" else f"{n:.1f} {unit}"
        n /= 1024.0
    return f"{n}"


def _table(rows: list[list], header: list[str], align: list[str], margin: int = 0) -> str:
    """``anemoi.utils.text.table``: plain-text table with a header rule."""
    def fmt(v):
        if isinstance(v, float):
            return f"{v:g}"
        return str(v)

    cells = [[fmt(v) for v in row] for row in [header] + rows]
    widths = [max(len(r[i]) for r in cells) for i in range(len(header))]
    pad = " " * margin
    lines = []
    for n, row in enumerate(cells):
        lines.append(pad + " │ ".join(f"{c:{a}{w}}" for c, a, w in zip(row, align, widths)))
        if n == 0:
            lines.append(pad + "─┼─".join("─" * w for w in widths))
    return "\n".join(lines)


def load_graph(path: Union[str, Path]):
    """``torch.load`` of a saved graph.  The file is a pickled ``HeteroData`` (create.py:158), which torch >= 2.6
    only loads with ``weights_only=False`` (the reference's own ``torch.load(path)`` fails there)."""
    return torch.load(path, map_location="cpu", weights_only=False)


class GraphDescriptor:
    """Class for descripting the graph."""

    def __init__(self, path: Union[str, Path], **kwargs):
        self.path = path
        self.graph = load_graph(self.path)
        self._device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")

    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        return t.to(self._device, non_blocking=True)

    @property
    def total_size(self):
        """Total size of the tensors in the graph (in bytes)."""
        total_size = 0
        for store in chain(self.graph.node_stores, self.graph.edge_stores):
            for value in store.values():
                if isinstance(value, torch.Tensor):
                    total_size += value.numel() * value.element_size()
        return total_size

    def get_node_summary(self) -> list[list]:
        """Per node set: name, number of nodes, attribute names, total attribute dimension, min / max latitude and
        longitude in degrees (describe.py:41-76)."""
        node_summary = []
        for name, nodes in self.graph.node_items():
            attributes = nodes.node_attrs()
            attributes.remove("x")
            x = self._dev(nodes.x)
            lo, hi = x.min(dim=0).values.tolist(), x.max(dim=0).values.tolist()
            node_summary.append(
                [
                    name,
                    nodes.num_nodes,
                    ", ".join(attributes),
                    sum(nodes[attr].shape[1] for attr in attributes if isinstance(nodes[attr], torch.Tensor)),
                    lo[0] / 2 / math.pi * 360,
                    hi[0] / 2 / math.pi * 360,
                    lo[1] / 2 / math.pi * 360,
                    hi[1] / 2 / math.pi * 360,
                ]
            )
        return node_summary

    def _isolated(self, row: torch.Tensor, num_nodes: int) -> int:
        """``num_nodes - len(torch.unique(row))`` without the sort: mark the endpoints that occur."""
        seen = torch.zeros(num_nodes, dtype=torch.bool, device=self._device)
        seen[self._dev(row).long()] = True
        return int(num_nodes - int(seen.sum().item()))

    def get_edge_summary(self) -> list[list]:
        """Per edge set: source, target, number of edges, isolated sources / targets, attribute dimension, attribute
        names (describe.py:78-105)."""
        edge_summary = []
        for (src_name, _, dst_name), edges in self.graph.edge_items():
            attributes = [a for a in edges.edge_attrs() if a != "edge_index"]
            edge_summary.append(
                [
                    src_name,
                    dst_name,
                    edges.num_edges,
                    self._isolated(edges.edge_index[0], self.graph[src_name].num_nodes),
                    self._isolated(edges.edge_index[1], self.graph[dst_name].num_nodes),
                    sum(edges[attr].shape[1] for attr in attributes),
                    ", ".join([f"{attr}({edges[attr].shape[1]}D)" for attr in attributes]),
                ]
            )
        return edge_summary

    def _stats(self, t: torch.Tensor) -> list[float]:
        v = self._dev(t).float()
        return torch.stack([v.min(), v.mean(), v.max(), v.std()]).tolist()  # one read-back per attribute

    def get_node_attribute_table(self) -> list[list]:
        node_attributes = []
        for node_name, node_store in self.graph.node_items():
            node_attr_names = node_store.node_attrs()
            node_attr_names.remove("x")  # Remove the coordinates from statistics table
            for node_attr_name in node_attr_names:
                node_attributes.append(
                    ["Node", node_name, node_attr_name, node_store[node_attr_name].dtype]
                    + self._stats(node_store[node_attr_name])
                )
        return node_attributes

    def get_edge_attribute_table(self) -> list[list]:
        edge_attributes = []
        for (source_name, _, target_name), edge_store in self.graph.edge_items():
            edge_attr_names = [a for a in edge_store.edge_attrs() if a != "edge_index"]  # not in the statistics table
            for edge_attr_name in edge_attr_names:
                edge_attributes.append(
                    ["Edge", f"{source_name}-->{target_name}", edge_attr_name, edge_store[edge_attr_name].dtype]
                    + self._stats(edge_store[edge_attr_name])
                )
        return edge_attributes

    def get_attribute_table(self) -> list[list]:
        """Get a table with the attributes of the graph."""
        attribute_table = []
        attribute_table.extend(self.get_node_attribute_table())
        attribute_table.extend(self.get_edge_attribute_table())
        return attribute_table

    def describe(self, show_attribute_distributions: Optional[bool] = True) -> None:
        """Describe the graph."""
        print()
        print(f"📦 Path       : {self.path}")
        print(f"💽 Size       : {_bytes(self.total_size)} ({self.total_size})")
        print()
        print("🪩  Nodes summary")
        print()
        print(
            _table(
                self.get_node_summary(),
                header=["Nodes name", "Num. nodes", "Attributes", "Attribute dim", "Min. latitude", "Max. latitude",
                        "Min. longitude", "Max. longitude"],
                align=["<", ">", ">", ">", ">", ">", ">", ">"],
                margin=3,
            )
        )  # fmt: skip
        print()
        print()
        print("🌐  Edges summary")
        print()
        print(
            _table(
                self.get_edge_summary(),
                header=["Source", "Target", "Num. edges", "Isolated Source", "Isolated Target", "Attribute dim",
                        "Attributes"],
                align=["<", "<", ">", ">", ">", ">", ">"],
                margin=3,
            )
        )  # fmt: skip
        print()
        if show_attribute_distributions:
            print()
            print("📊 Attribute distributions")
            print()
            print(
                _table(
                    self.get_attribute_table(),
                    header=["Type", "Source", "Name", "Dtype", "Min.", "Mean", "Max.", "Std. dev."],
                    align=["<", "<", ">", ">", ">", ">", ">", ">"],
                    margin=3,
                )
            )
            print()
        print("🔋 Graph ready.")
        print()
