"""B200-native edge-construction path of anemoi-graphs behind the reference's builder API.

Module layout mirrors ``anemoi.graphs`` for the hot path, so a recipe's ``_target_`` strings
(``anemoi.graphs.edges.KNNEdges`` ...) resolve here (``config.resolve_target``):

    create.GraphCreator                         /root/reference/src/anemoi/graphs/create.py
    edges.{KNNEdges,CutOffEdges,MultiScaleEdges} .../edges/builder.py
    edges.attributes.{EdgeLength,EdgeDirection}  .../edges/attributes.py
    nodes.{TriNodes,LimitedAreaTriNodes,StretchedTriNodes,LatLonNodes,NPZFileNodes,...}
    generate.masks.KNNAreaMaskBuilder            .../generate/masks.py

All arithmetic runs in hand-written sm_100a CUDA kernels reached through the C ABI of
``lib/libagx_b200.so`` (``include/agx_b200.h``).  There is no CPU fallback.
"""

EARTH_RADIUS = 6371.0  # km; /root/reference/src/anemoi/graphs/__init__.py

__version__ = "0.1.0"
