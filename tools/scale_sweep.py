"""BASELINE.json config 5 across GPUs (run under torchrun on the GPU box): KNN / cut-off edges between 1 M uniform
reference points and 100 M uniform query points through the builder API, query nodes sharded over the ranks
(``device.shard_world``: 100 M >= AGX_SHARD_MIN_QUERIES), per-rank blocks all-gathered so that every rank ends
with the complete edge list (``device.ChunkedGather``: every finished chunk of the search is all-gathered on a
second stream while the next one is searched).  Times are CUDA events around ``compute_edge_index`` + the wait for the gathers, max over
ranks; a second column gives the time without the all-gather (each rank keeps its block).

    python -m torch.distributed.run --nproc-per-node N tools/scale_sweep.py [--queries 100000000] [--ks 3,16]
"""
import argparse
import json
import os
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist

from anemoi_graphs_b200 import device as agx_device
from anemoi_graphs_b200 import grids, ops
from anemoi_graphs_b200.edges import CutOffEdges, KNNEdges
from anemoi_graphs_b200.graph import HeteroData


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refs", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--ks", default="3,16")
    ap.add_argument("--degrees", default="8")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    agx_device.set_resident(True)
    graph = HeteroData()
    graph["ref"].x = grids.latlon_deg_to_x(*grids.uniform_sphere(args.refs, seed=1234)).cuda()
    graph["ref"].node_type = "LatLonNodes"
    graph["q"].x = grids.latlon_deg_to_x(*grids.uniform_sphere(args.queries, seed=4321)).cuda()
    graph["q"].node_type = "LatLonNodes"

    def timed(builder, gather: bool):
        times = []
        n_edges = 0
        for rep in range(args.reps + 1):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if gather:
                ei = builder.get_edge_index_device(graph)
                agx_device.wait_for(ei)
            else:  # the rank's own block only: the search without the exchange
                saved = agx_device.ChunkedGather.chunk_done
                agx_device.ChunkedGather.chunk_done = lambda self, c, ready=None: None
                try:
                    ei = builder.get_edge_index_device(graph)
                finally:
                    agx_device.ChunkedGather.chunk_done = saved
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            if rep > 0:
                times.append(ms)
            n_edges = int(ei.shape[1])
            del ei
        return float(np.median(times)), n_edges

    for k in [int(v) for v in args.ks.split(",")]:
        b = KNNEdges("ref", "q", k)
        ms, e = timed(b, True)
        ms_local, _ = timed(b, False)
        if rank == 0:
            print(json.dumps(dict(op="knn", n_gpus=world, n_ref=args.refs, n_query=args.queries, k=k, edges=e, ms=round(ms, 2),
                                  edges_per_s=round(e / ms * 1e3), ms_without_all_gather=round(ms_local, 2))), flush=True)  # fmt: skip
    for deg in [int(v) for v in args.degrees.split(",")]:
        b = CutOffEdges("ref", "q", 1.0)
        b.get_cutoff_radius = lambda graph, mask_attr=None, r=float(np.arccos(1.0 - 2.0 * deg / args.refs)): r
        ms, e = timed(b, True)
        ms_local, _ = timed(b, False)
        if rank == 0:
            print(json.dumps(dict(op="cutoff", n_gpus=world, n_ref=args.refs, n_query=args.queries, mean_degree=round(e / args.queries, 2),
                                  edges=e, ms=round(ms, 2), edges_per_s=round(e / ms * 1e3), ms_without_all_gather=round(ms_local, 2))), flush=True)  # fmt: skip
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
