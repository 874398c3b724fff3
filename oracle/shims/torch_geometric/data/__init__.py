from .storage import EdgeStorage, NodeStorage  # noqa: F401


class HeteroData:
    def __init__(self):
        self._nodes = {}
        self._edges = {}

    def __getitem__(self, key):
        if isinstance(key, tuple):
            return self._edges.setdefault(tuple(key), EdgeStorage())
        return self._nodes.setdefault(key, NodeStorage())

    def __contains__(self, key):
        return key in self._nodes or key in self._edges

    @property
    def node_types(self):
        return list(self._nodes)

    @property
    def edge_types(self):
        return list(self._edges)

    @property
    def node_stores(self):
        return list(self._nodes.values())

    @property
    def edge_stores(self):
        return list(self._edges.values())

    def node_items(self):
        return list(self._nodes.items())

    def edge_items(self):
        return list(self._edges.items())

    @property
    def num_nodes(self):
        return sum(s.num_nodes for s in self._nodes.values())

    @property
    def num_edges(self):
        return sum(s.num_edges for s in self._edges.values())
