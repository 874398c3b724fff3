"""Graph container used by the builders.

The reference stores everything in ``torch_geometric.data.HeteroData``
(/root/reference/src/anemoi/graphs/create.py:184, edges/builder.py:89-115).  When
``torch_geometric`` is importable the real class is used unchanged, so a graph written by
this package is the same pickle a reference user would get.  When it is not (this image),
``HeteroData`` below provides the subset of the API the hot path touches: string keys give
node stores, ``(src, "to", dst)`` tuples give edge stores, stores behave like attribute
dictionaries.
"""

from __future__ import annotations

from typing import Any, Iterator

import torch

try:  # pragma: no cover - depends on the environment
    from torch_geometric.data import HeteroData as _PygHeteroData  # type: ignore
    from torch_geometric.data.storage import EdgeStorage as _PygEdgeStorage  # type: ignore
    from torch_geometric.data.storage import NodeStorage as _PygNodeStorage  # type: ignore

    HAVE_PYG = True
except Exception:  # ModuleNotFoundError in this image
    HAVE_PYG = False


class _Storage:
    """Attribute dictionary with the access patterns the builders use."""

    def __init__(self, key=None) -> None:
        object.__setattr__(self, "_mapping", {})
        object.__setattr__(self, "_key", key)

    # mapping protocol -------------------------------------------------------------------
    def __getitem__(self, name: str) -> Any:
        return self._mapping[name]

    def __setitem__(self, name: str, value: Any) -> None:
        self._mapping[name] = value

    def __delitem__(self, name: str) -> None:
        del self._mapping[name]

    def __contains__(self, name: str) -> bool:
        return name in self._mapping

    def __iter__(self) -> Iterator[str]:
        return iter(list(self._mapping.keys()))

    def __len__(self) -> int:
        return len(self._mapping)

    def keys(self):
        return list(self._mapping.keys())

    def values(self):
        return list(self._mapping.values())

    def items(self):
        return list(self._mapping.items())

    def get(self, name: str, default: Any = None) -> Any:
        return self._mapping.get(name, default)

    # attribute protocol -----------------------------------------------------------------
    def __getattr__(self, name: str) -> Any:
        if name.startswith("__"):
            raise AttributeError(name)
        try:
            return object.__getattribute__(self, "_mapping")[name]
        except KeyError:
            raise AttributeError(f"'{type(self).__name__}' has no attribute '{name}'") from None

    def __setattr__(self, name: str, value: Any) -> None:
        self._mapping[name] = value

    def __delattr__(self, name: str) -> None:
        try:
            del self._mapping[name]
        except KeyError:
            raise AttributeError(name) from None

    def __getstate__(self):
        return {"_mapping": self._mapping, "_key": self._key}

    def __setstate__(self, state):
        object.__setattr__(self, "_mapping", state["_mapping"])
        object.__setattr__(self, "_key", state["_key"])

    def to_dict(self) -> dict:
        return dict(self._mapping)

    def __repr__(self) -> str:
        parts = []
        for k, v in self._mapping.items():
            if isinstance(v, torch.Tensor):
                parts.append(f"{k}={list(v.shape)}")
            else:
                parts.append(f"{k}={type(v).__name__}" if not isinstance(v, (str, int, float)) else f"{k}={v!r}")
        return "{" + ", ".join(parts) + "}"


class _NodeStorage(_Storage):
    @property
    def num_nodes(self) -> int:
        x = self._mapping.get("x")
        return int(x.shape[0]) if x is not None else 0

    def node_attrs(self) -> list[str]:
        n = self.num_nodes
        return [k for k, v in self._mapping.items() if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n]


class _EdgeStorage(_Storage):
    @property
    def num_edges(self) -> int:
        ei = self._mapping.get("edge_index")
        return int(ei.shape[1]) if ei is not None else 0

    def edge_attrs(self) -> list[str]:
        """Tensor attributes with one entry per edge - ``edge_index`` included, as in torch_geometric (the reference
        removes it by name, describe.py:92)."""
        e = self.num_edges
        return [
            k
            for k, v in self._mapping.items()
            if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[-1 if k == "edge_index" else 0] == e
        ]


class _HeteroData:
    def __init__(self) -> None:
        self._node_store: dict[str, _NodeStorage] = {}
        self._edge_store: dict[tuple, _EdgeStorage] = {}

    def __getitem__(self, key):
        if isinstance(key, tuple):
            key = tuple(key)
            if key not in self._edge_store:
                self._edge_store[key] = _EdgeStorage(key)
            return self._edge_store[key]
        if key not in self._node_store:
            self._node_store[key] = _NodeStorage(key)
        return self._node_store[key]

    def __contains__(self, key) -> bool:
        return key in self._node_store or (isinstance(key, tuple) and tuple(key) in self._edge_store)

    def __delitem__(self, key) -> None:
        if isinstance(key, tuple):
            del self._edge_store[tuple(key)]
        else:
            del self._node_store[key]

    @property
    def node_types(self) -> list[str]:
        return list(self._node_store.keys())

    @property
    def edge_types(self) -> list[tuple]:
        return list(self._edge_store.keys())

    @property
    def node_stores(self) -> list[_NodeStorage]:
        return list(self._node_store.values())

    @property
    def edge_stores(self) -> list[_EdgeStorage]:
        return list(self._edge_store.values())

    def node_items(self):
        return list(self._node_store.items())

    def edge_items(self):
        return list(self._edge_store.items())

    @property
    def num_nodes(self) -> int:
        return sum(s.num_nodes for s in self._node_store.values())

    @property
    def num_edges(self) -> int:
        return sum(s.num_edges for s in self._edge_store.values())

    def to_dict(self) -> dict:
        out = {k: v.to_dict() for k, v in self._node_store.items()}
        out.update({k: v.to_dict() for k, v in self._edge_store.items()})
        return out

    def __repr__(self) -> str:
        lines = ["HeteroData("]
        for k, v in self._node_store.items():
            lines.append(f"  {k}={v!r},")
        for k, v in self._edge_store.items():
            lines.append(f"  {k}={v!r},")
        lines.append(")")
        return "\n".join(lines)


if HAVE_PYG:  # pragma: no cover
    HeteroData = _PygHeteroData
    NodeStorage = _PygNodeStorage
    EdgeStorage = _PygEdgeStorage
else:
    HeteroData = _HeteroData
    NodeStorage = _NodeStorage
    EdgeStorage = _EdgeStorage
    # ``torch.load`` (weights_only=True, default since torch 2.6) refuses unknown classes;
    # register ours so ``torch.load(path)`` of a saved graph keeps working
    # (/root/reference/src/anemoi/graphs/describe.py:26 uses the bare call).
    try:
        torch.serialization.add_safe_globals([_HeteroData, _NodeStorage, _EdgeStorage, _Storage])
    except Exception:  # pragma: no cover
        pass
