"""Edge attributes (/root/reference/src/anemoi/graphs/edges/attributes.py:24-157).

``EdgeLength`` and ``EdgeDirection`` keep the reference's constructor arguments and ``compute(graph,
edges_name)`` contract (float32 ``(E, 1)`` / ``(E, 2)`` tensors, normalised over the whole edge set).
Raw values, the global reduction for the normalisation and the final float32 cast are one fused CUDA
kernel pair (``ops.edge_attributes``); when both attributes are requested for an edge set
(``compute_attributes``) they share a single pass over ``edge_index``.
"""

from __future__ import annotations

from abc import ABC

import torch

from .. import device as _device
from .. import ops
from .._cabi import NORM_CODES


def _check_norm(norm) -> None:
    if norm not in NORM_CODES:
        # normalise.py:53-55
        raise ValueError(
            f"Attribute normalisation \"{norm}\" is not valid. Options are: 'l1', 'l2', 'unit-max' or 'unit-std'."
        )


class BaseEdgeAttribute(ABC):
    """Base class for edge attributes."""

    _agx_device_aware = True

    def __init__(self, norm: str | None = None) -> None:
        self.norm = norm

    def _kernel_args(self) -> dict:
        raise NotImplementedError

    def _pick(self, length, direction) -> torch.Tensor:
        raise NotImplementedError

    def compute(self, graph, edges_name: tuple[str, str, str], *args, **kwargs) -> torch.Tensor:
        """Compute the edge attributes."""
        out = compute_attributes(graph, edges_name, {"value": self})["value"]
        _device.maybe_flush()
        return out


class EdgeDirection(BaseEdgeAttribute):
    """Edge direction feature (edges/attributes.py:56-104).

    1. Rotated features: the target nodes are rotated to the north pole to compute the edge direction.
    2. Non-rotated features: difference in latitude and longitude between the source and target nodes.

    Attributes
    ----------
    norm : Optional[str]
        Normalisation method. Options: None, "l1", "l2", "unit-max", "unit-range", "unit-std".
    luse_rotated_features : bool
        Whether to use rotated features.
    """

    def __init__(self, norm: str | None = None, luse_rotated_features: bool = True) -> None:
        super().__init__(norm)
        self.luse_rotated_features = luse_rotated_features

    def _kernel_args(self) -> dict:
        return dict(direction=True, direction_norm=self.norm, direction_rotated=bool(self.luse_rotated_features))

    def _pick(self, length, direction):
        return direction


class EdgeLength(BaseEdgeAttribute):
    """Edge length feature: haversine distance on the unit sphere (edges/attributes.py:107-157).

    Attributes
    ----------
    norm : Optional[str]
        Normalisation method. Options: None, "l1", "l2", "unit-max", "unit-range", "unit-std".
    invert : bool
        Whether to invert the edge lengths, i.e. 1 - edge_length. Defaults to False.
    """

    def __init__(self, norm: str | None = None, invert: bool = False) -> None:
        super().__init__(norm)
        self.invert = invert

    def _kernel_args(self) -> dict:
        return dict(length=True, length_norm=self.norm, length_invert=bool(self.invert))

    def _pick(self, length, direction):
        return length


def _check_nodes(graph, edges_name) -> tuple[str, str]:
    source_name, _, target_name = edges_name
    assert (
        source_name in graph.node_types
    ), f"Node \"{source_name}\" not found in graph. Optional nodes are {', '.join(graph.node_types)}."
    assert (
        target_name in graph.node_types
    ), f"Node \"{target_name}\" not found in graph. Optional nodes are {', '.join(graph.node_types)}."
    return source_name, target_name


def compute_attributes(graph, edges_name: tuple[str, str, str], attrs: dict) -> dict:
    """Evaluate a set of attribute objects on one edge set.

    Pairs one ``EdgeLength`` with one ``EdgeDirection`` per kernel pass; foreign attribute objects (anything
    with a ``compute(graph, edges_name)`` method) are called as the reference would call them
    (edges/builder.py:133)."""
    ours = {k: a for k, a in attrs.items() if isinstance(a, (EdgeLength, EdgeDirection))}
    out: dict = {}
    for k, a in attrs.items():
        if k not in ours:
            # a foreign (reference-style plugin) attribute may read ``x`` / ``edge_index`` on the host: give every
            # provisional node set its final order and wait for the pending device->host copies first
            _device.flush()
            out[k] = a.compute(graph, edges_name)
    if not ours:
        return out
    source_name, target_name = _check_nodes(graph, edges_name)
    for a in ours.values():
        _check_norm(a.norm)
    store = graph[tuple(edges_name)]
    edge_index = _device.device_edge_index(store)
    host_side = not store["edge_index"].is_cuda
    # A row of the edge list may still be in the provisional numbering of its node set (device.Provisional): the
    # attributes depend on coordinates only, so they are evaluated right away against the matching (provisional)
    # node records.  Anything else - final rows against a provisional node set - needs the final order first.
    meta = _device.edge_meta(edge_index)
    # KNN edges whose index-order ties are re-decided when the node order resolves: evaluated NOW (the trigonometry
    # runs in the shadow of the host sort) with the re-decided targets' edges kept out of the statistics; those edges
    # are evaluated again after the re-decision and the normalisation is applied then (``_deferred_attributes``)
    pending = meta.fixup if meta is not None else None
    if pending is not None and pending.done:
        pending = None
    _, w = _device.world()
    if pending is not None and ((w > 1 and not _device.sharded_output()) or meta.tie_flags is None or len(ours) > 2):
        pending.resolve()
        pending = None
    tags = _device.row_tags(edge_index)
    for row, name in ((0, source_name), (1, target_name)):
        prov = _device.active_provisional(graph[name])
        if prov is not None and tags[row] is not prov:
            prov.resolve()
    if source_name == target_name and tags[0] is not tags[1]:
        for prov in tags:
            if prov is not None:
                prov.resolve()
    src = _device.node_tables(graph[source_name], provisional_ok=True)
    dst = _device.node_tables(graph[target_name], provisional_ok=True)
    lengths = [(k, a) for k, a in ours.items() if isinstance(a, EdgeLength)]
    dirs = [(k, a) for k, a in ours.items() if isinstance(a, EdgeDirection)]
    if pending is not None and len(lengths) <= 1 and len(dirs) <= 1:
        out.update(_deferred_attributes(pending, meta, edge_index, src, dst, lengths, dirs, host_side))
        return {k: out[k] for k in attrs}
    if pending is not None:
        pending.resolve()
        src = _device.node_tables(graph[source_name], provisional_ok=True)
        dst = _device.node_tables(graph[target_name], provisional_ok=True)
    local = meta.local if meta is not None else None  # set by a sharded builder: this rank's own columns
    while lengths or dirs:
        kl = lengths.pop(0) if lengths else None
        kd = dirs.pop(0) if dirs else None
        shard = meta.shard if (meta is not None and _device.sharded_output()) else None
        args = dict(length=False, direction=False, sharded=w > 1, local=local, shard=shard,
                    regular_k=meta.regular_k if meta is not None else 0)
        if kl:
            args.update(kl[1]._kernel_args())
        if kd:
            args.update(kd[1]._kernel_args())
        ln, dr = ops.edge_attributes(edge_index, src, dst, **args)
        if kl:
            out[kl[0]] = _device.to_host(ln, shard) if host_side else ln
        if kd:
            out[kd[0]] = _device.to_host(dr, shard) if host_side else dr
    return {k: out[k] for k in attrs}  # the recipe's order


def _deferred_attributes(prov, meta, edge_index, src, dst, lengths, dirs, host_side: bool) -> dict:
    """One EdgeLength and/or one EdgeDirection of an edge set that still waits for a KNN tie re-decision: raw pass now,
    the re-decided edges again when the order resolves (a fixup: still provisional numbering, the same node records),
    scaling and host copies after that (a finalizer).  Returns the tensors the graph stores: device tensors that are
    complete in stream order, or pinned host tensors that are complete at ``flush()``."""
    args = dict(length=False, direction=False)
    for _, a in lengths + dirs:
        args.update(a._kernel_args())
    flag_list, flag_count = meta.tie_list if meta.tie_list is not None else (None, None)
    job = ops.DeferredEdgeAttributes(
        edge_index, src, dst, meta.tie_flags, flag_list=flag_list, flag_count=flag_count, regular_k=meta.regular_k,
        flag_base=meta.flag_base, shard=meta.shard, **args
    )
    job.raw()
    prov.add_fixup(lambda p, job=job: job.patch())  # registered after the builder's re-decision: runs after it
    results, host_targets = {}, []
    for name, _ in lengths:
        results[name] = job.out_len
    for name, _ in dirs:
        results[name] = job.out_dir
    if host_side:
        shard = meta.shard if (meta.shard is not None and _device.sharded_output()) else None
        for name, dev_out in list(results.items()):
            if shard is None:
                host = dest = torch.empty(dev_out.shape, dtype=dev_out.dtype, pin_memory=True)
            else:  # the complete array in the shared host buffer; this rank fills its rows
                host = _device.host_tensor((shard.total, int(dev_out.shape[1])), dev_out.dtype, require_shared=True)
                dest = host[shard.offset : shard.offset + int(dev_out.shape[0])]
            host_targets.append((dev_out, dest))
            results[name] = host

    def finalize(p, job=job, host_targets=host_targets):
        job.apply()
        for dev_out, dest in host_targets:
            _device.to_host_into(dev_out, dest)

    prov.add_finalizer(finalize)
    return results
