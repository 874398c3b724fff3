"""BASELINE.json config 5 (run on the GPU box, alone or under torchrun): KNN / cut-off edges between 1 M uniform reference
points and 1 M / 10 M / 100 M uniform query points through the builder API, k in {3, 8, 16, 32}, cut-off radii for mean
degree 8 / 16 / 64, query nodes sharded over the ranks.

Per row two multi-GPU forms are timed (CUDA events around ``get_edge_index_device`` + the wait for any exchange, max over
ranks, median of ``--reps``):

* ``ms_sharded``  - sharded OUTPUT mode: every rank keeps the block of its own queries (no edge data leaves the GPU);
* ``ms_gathered`` - every rank ends with the complete list (chunked all-gathers overlapping the search).

``--sklearn`` (single process only) times the reference's own call beside every 1 M-query row:
``NearestNeighbors(metric="haversine", n_jobs=4).kneighbors_graph / radius_neighbors_graph`` (edges/builder.py:259-265,
364-366) on the same points, fit included, and checks the edge sets against it on a 20 000-query sample.

    python tools/scale_sweep.py --sklearn                       # N = 1
    python -m torch.distributed.run --nproc-per-node N tools/scale_sweep.py
"""
import argparse
import json
import os
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist

from anemoi_graphs_b200 import device as agx_device
from anemoi_graphs_b200 import grids
from anemoi_graphs_b200.edges import CutOffEdges, KNNEdges
from anemoi_graphs_b200.graph import HeteroData


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refs", type=int, default=1_000_000)
    ap.add_argument("--queries", default="1000000,10000000,100000000")
    ap.add_argument("--ks", default="3,8,16,32")
    ap.add_argument("--degrees", default="8,16,64")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--sklearn", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    agx_device.set_resident(True)
    sizes = [int(v) for v in args.queries.split(",")]
    ref_host = grids.latlon_deg_to_x(*grids.uniform_sphere(args.refs, seed=1234))
    q_host_all = grids.latlon_deg_to_x(*grids.uniform_sphere(max(sizes), seed=4321))

    def emit(row):
        if rank == 0:
            print(json.dumps(row), flush=True)

    def timed(builder, graph, sharded: bool):
        prev = agx_device.set_sharded_output(sharded)
        times, n_local = [], 0
        try:
            for rep in range(args.reps + 1):
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ei = builder.get_edge_index_device(graph)
                agx_device.wait_for(ei)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b)
                if world > 1:
                    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                if rep > 0:
                    times.append(ms)
                n_local = int(ei.shape[1])
                shard = agx_device.edge_shard(ei)
                total = shard.total if (shard is not None and sharded and world > 1) else n_local
                del ei
        finally:
            agx_device.set_sharded_output(prev)
        return float(np.median(times)), total

    for nq in sizes:
        graph = HeteroData()
        graph["ref"].x = ref_host.cuda()
        graph["ref"].node_type = "LatLonNodes"
        graph["q"].x = q_host_all[:nq].cuda()
        graph["q"].node_type = "LatLonNodes"
        builders = [("knn", k, KNNEdges("ref", "q", k)) for k in (int(v) for v in args.ks.split(","))]
        for deg in (int(v) for v in args.degrees.split(",")):
            b = CutOffEdges("ref", "q", 1.0)
            b.get_cutoff_radius = lambda graph, mask_attr=None, r=float(np.arccos(1.0 - 2.0 * deg / args.refs)): r
            builders.append(("cutoff", deg, b))
        for op, param, b in builders:
            row = dict(op=op, n_gpus=world, n_ref=args.refs, n_query=nq)
            row["k" if op == "knn" else "target_degree"] = param
            ms_sh, edges = timed(b, graph, True)
            row.update(edges=edges, ms_sharded=round(ms_sh, 3), edges_per_s_sharded=round(edges / ms_sh * 1e3))
            if op == "cutoff":
                row["mean_degree"] = round(edges / nq, 2)
                row["radius_rad"] = round(b.get_cutoff_radius(graph), 6)
            if world > 1:
                ms_g, _ = timed(b, graph, False)
                row.update(ms_gathered=round(ms_g, 3), edges_per_s_gathered=round(edges / ms_g * 1e3))
            if args.sklearn and world == 1 and nq == 1_000_000:
                from sklearn.neighbors import NearestNeighbors

                rx, qx = ref_host.numpy(), q_host_all[:nq].numpy()
                t0 = time.perf_counter()
                nn = NearestNeighbors(metric="haversine", n_jobs=4).fit(rx)
                if op == "knn":
                    adj = nn.kneighbors_graph(qx, n_neighbors=param, mode="distance").tocoo()
                else:
                    adj = nn.radius_neighbors_graph(qx, radius=b.get_cutoff_radius(graph)).tocoo()
                t_sk = time.perf_counter() - t0
                row.update(sklearn_s=round(t_sk, 2), sklearn_threads=4, host_cpus=os.cpu_count(), sklearn_edges=int(adj.nnz),
                           speedup_vs_sklearn=round(t_sk * 1e3 / ms_sh, 1))  # fmt: skip
                # parity on the first 20 000 queries (uniform random points: no ties)
                ei = b.get_edge_index_device(graph).cpu().numpy()
                m = 20_000
                got = ei[:, ei[1] < m]
                got = got[:, np.lexsort((got[0], got[1]))]
                ref = np.stack([adj.col[adj.row < m], adj.row[adj.row < m]]).astype(np.int32)
                ref = ref[:, np.lexsort((ref[0], ref[1]))]
                row["matches_sklearn_on_sample"] = bool(got.shape == ref.shape and np.array_equal(got, ref))
                row["edges_match_sklearn_count"] = bool(int(adj.nnz) == edges)
            emit(row)
        del graph
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
