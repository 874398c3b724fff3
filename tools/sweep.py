"""BASELINE.json config 5: neighbour-search scaling sweep on a uniform random sphere (run on the GPU box).

    python tools/sweep.py [--refs 1000000] [--queries 1000000,10000000] [--ks 3,8,16,32] [--degrees 8,16,64]
                          [--sorted] [--json gpurun_out/sweep.json]

Reference and query points are `grids.uniform_sphere` (lat = arcsin U(-1, 1), lon = U(0, 2 pi), float32,
default_rng seeds 1234 / 4321).  Times are CUDA-event averages of the C-ABI calls with inputs resident in HBM:
KNN = `agx_knn` alone (the index build is timed separately); cut-off = count + scan + fill.
`--sorted` additionally times the queries pre-sorted by latitude band and longitude (a coherent order).
"""
import argparse
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from anemoi_graphs_b200 import grids, ops


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refs", type=int, default=1_000_000)
    ap.add_argument("--queries", default="1000000,10000000")
    ap.add_argument("--ks", default="3,8,16,32")
    ap.add_argument("--degrees", default="8,16,64")
    ap.add_argument("--sorted", action="store_true")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    ref = grids.latlon_deg_to_x(*grids.uniform_sphere(args.refs, seed=1234)).cuda()
    rows = []
    for nq in [int(v) for v in args.queries.split(",")]:
        q = grids.latlon_deg_to_x(*grids.uniform_sphere(nq, seed=4321)).cuda()
        variants = [("random", q)]
        if args.sorted:
            qn = q.cpu().numpy()
            band = np.floor((qn[:, 0] + np.pi / 2) / (np.pi / 2048)).astype(np.int64)
            variants.append(("lat-band sorted", q[torch.from_numpy(np.lexsort((qn[:, 1], band))).cuda()]))
        for tag, qq in variants:
            for k in [int(v) for v in args.ks.split(",")]:
                t_build = timed(lambda: ops.NeighbourIndex(ref, hint_k=k).close())
                with ops.NeighbourIndex(ref, hint_k=k) as ix:
                    out = torch.empty((2, nq * k), dtype=torch.int32, device="cuda")
                    ms = timed(lambda: ix.knn(qq, k, out=out))
                    st = ops.new_stats("cuda")
                    ix.knn(qq, k, out=out, stats=st)
                    st = st.cpu().tolist()
                row = dict(op="knn", order=tag, n_ref=args.refs, n_query=nq, k=k, ms=round(ms, 3), index_build_ms=round(t_build, 3),
                           edges_per_s=round(nq * k / ms * 1e3), queries_per_s=round(nq / ms * 1e3),
                           f64_refined=st[0], tied=st[1], widened=st[2], staged_per_tile=round(st[3] / 32 / ((nq + 31) // 32), 1))  # fmt: skip
                rows.append(row)
                print(json.dumps(row), flush=True)
            for deg in [int(v) for v in args.degrees.split(",")]:
                # mean degree = n_ref * cap area / sphere area = n_ref * (1 - cos r) / 2
                r = float(np.arccos(1.0 - 2.0 * deg / args.refs))
                with ops.NeighbourIndex(ref, hint_radius=r) as ix:
                    offsets, e = ix.radius_count(qq, r)
                    out = torch.empty((2, e), dtype=torch.int32, device="cuda")  # allocation is not part of the search

                    def run():
                        off, tot = ix.radius_count(qq, r)
                        ix.radius_fill(qq, r, off, tot, out)

                    ms = timed(run)
                row = dict(op="cutoff", order=tag, n_ref=args.refs, n_query=nq, radius=r, mean_degree=round(e / nq, 2),
                           ms=round(ms, 3), edges=e, edges_per_s=round(e / ms * 1e3))  # fmt: skip
                rows.append(row)
                print(json.dumps(row), flush=True)
                del out
        del q
    if args.json:
        pathlib.Path(args.json).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
