import torch


class BaseStorage:
    def __init__(self):
        object.__setattr__(self, "_d", {})

    def __getitem__(self, k):
        return self._d[k]

    def __setitem__(self, k, v):
        self._d[k] = v

    def __delitem__(self, k):
        del self._d[k]

    def __contains__(self, k):
        return k in self._d

    def __iter__(self):
        return iter(list(self._d))

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        try:
            return object.__getattribute__(self, "_d")[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self._d[k] = v

    def __getstate__(self):
        return self._d

    def __setstate__(self, s):
        object.__setattr__(self, "_d", s)

    def get(self, k, default=None):
        return self._d.get(k, default)

    def keys(self):
        return list(self._d)

    def items(self):
        return list(self._d.items())


class NodeStorage(BaseStorage):
    @property
    def num_nodes(self):
        return int(self._d["x"].shape[0]) if "x" in self._d else 0

    def node_attrs(self):
        n = self.num_nodes
        return [k for k, v in self._d.items() if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n]


class EdgeStorage(BaseStorage):
    @property
    def num_edges(self):
        return int(self._d["edge_index"].shape[1]) if "edge_index" in self._d else 0

    def edge_attrs(self):
        e = self.num_edges
        return [
            k
            for k, v in self._d.items()
            if isinstance(v, torch.Tensor) and k != "edge_index" and v.dim() > 0 and v.shape[0] == e
        ]
