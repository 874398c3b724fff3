"""BASELINE.json config 4 at scale (run on the GPU box): a 2.5 km limited-area patch plus the global O96 grid as data
nodes, a stretched triangular hidden mesh (coarse globally, fine over the patch), KNN k=16 decoder, cut-off encoder
restricted to the patch (`source_mask_attr_name`), multi-scale processor.  Non-uniform density is the point
(SURVEY.md H5): stage times are printed and samples of every edge set are checked against the oracle.

    python tools/lam_check.py [--patch 1000] [--global-res 5] [--lam-res 10] [--k 16]
    python tools/lam_check.py --hidden hex --hex-res 5     # the same data nodes under a global H3 hidden mesh
"""
import argparse
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from anemoi_graphs_b200 import device as agx_device
from anemoi_graphs_b200 import grids
from anemoi_graphs_b200.edges import CutOffEdges, KNNEdges, MultiScaleEdges
from anemoi_graphs_b200.graph import HeteroData
from anemoi_graphs_b200.nodes import HexNodes, StretchedTriNodes
from oracle import ref_path as R


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patch", type=int, default=1000)
    ap.add_argument("--global-res", type=int, default=5)
    ap.add_argument("--lam-res", type=int, default=10)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--cutoff", type=float, default=0.6)
    ap.add_argument("--hidden", choices=["tri", "hex"], default="tri")
    ap.add_argument("--hex-res", type=int, default=5)
    args = ap.parse_args()
    lam_lat, lam_lon = grids.lam_patch(args.patch, args.patch, 2.5)
    g_lat, g_lon = grids.octahedral_grid(96)
    x = grids.latlon_deg_to_x(np.concatenate([lam_lat, g_lat]), np.concatenate([lam_lon, g_lon]))
    cutout = torch.zeros((x.shape[0], 1), dtype=torch.bool)
    cutout[: lam_lat.size] = True
    agx_device.set_resident(True)
    attrs = {
        "edge_length": {"_target_": "anemoi.graphs.edges.attributes.EdgeLength", "norm": "unit-std"},
        "edge_dirs": {"_target_": "anemoi.graphs.edges.attributes.EdgeDirection", "norm": "unit-std"},
    }
    for rep in range(2):
        graph = HeteroData()
        graph["data"].x = x.cuda()
        graph["data"].node_type = "LatLonNodes"
        graph["data"]["cutout"] = cutout.cuda()
        marks = []

        def mark(name):
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

        mark("start")
        if args.hidden == "hex":
            graph = HexNodes(args.hex_res, "hidden").update_graph(graph, {})
        else:
            graph = StretchedTriNodes(args.global_res, args.lam_res, "hidden", "data", "cutout", margin_radius_km=100.0).update_graph(graph, {})
        mark("hidden nodes")
        MultiScaleEdges("hidden", "hidden", 1).update_graph(graph, attrs)
        mark("multiscale + attrs")
        KNNEdges("hidden", "data", args.k).update_graph(graph, attrs)
        mark(f"knn k={args.k} + attrs")
        CutOffEdges("data", "hidden", args.cutoff, source_mask_attr_name="cutout").update_graph(graph, attrs)
        mark("cutoff (patch sources) + attrs")
        line = "  ".join(f"{n}={1e3 * (t - marks[i][1]):.2f}ms" for i, (n, t) in enumerate(marks[1:]))
        print(f"rep {rep}: data={x.shape[0]} hidden={graph['hidden'].x.shape[0]} "
              f"edges ms/knn/cut={graph['hidden', 'to', 'hidden'].edge_index.shape[1]}/"
              f"{graph['hidden', 'to', 'data'].edge_index.shape[1]}/{graph['data', 'to', 'hidden'].edge_index.shape[1]}  {line}", flush=True)  # fmt: skip
    # ---- oracle samples -------------------------------------------------------------------------------------
    dx, hx = x.numpy(), graph["hidden"].x.cpu().numpy()
    if args.hidden == "hex":
        ms = graph["hidden", "to", "hidden"].edge_index.cpu().numpy().astype(np.int64)
        want_n = sum(6 * (2 + 120 * 7**r) - 12 for r in range(args.hex_res + 1))
        fwd, bwd = np.sort(ms[0] * hx.shape[0] + ms[1]), np.sort(ms[1] * hx.shape[0] + ms[0])
        ok = ms.shape[1] == want_n and np.array_equal(fwd, bwd) and np.unique(fwd).size == fwd.size
        print("hex multiscale:", "OK" if ok else "MISMATCH", f"({ms.shape[1]} edges, H3 levels 0..{args.hex_res}: {want_n})")
    rng = np.random.default_rng(0)
    # KNN: a sample of data queries (half from the patch, half global)
    qs = np.sort(np.concatenate([rng.choice(lam_lat.size, 4000, replace=False), lam_lat.size + rng.choice(g_lat.size, 4000, replace=False)]))
    want, info = R.knn_edges_canonical(hx, dx[qs], args.k)
    ei = graph["hidden", "to", "data"].edge_index.cpu().numpy().reshape(2, -1, args.k)[:, qs, :]
    ei[1] = np.arange(qs.size)[:, None]
    got = R.canonical_sort(ei.reshape(2, -1))
    print("knn sample:", "OK" if np.array_equal(got, want) else "MISMATCH", f"({qs.size} queries, {info['tied_queries'].size} tied)")
    # cut-off: a sample of hidden targets against the patch sources
    radius = R.cutoff_radius(hx, args.cutoff)
    ts = np.sort(rng.choice(hx.shape[0], 300, replace=False))
    src_sel = np.arange(lam_lat.size)
    want = R.cutoff_edges(dx[src_sel], hx[ts], args.cutoff, radius=radius)
    want = R.canonical_sort(np.stack([src_sel[want[0]], ts[want[1]]]).astype(np.int32))
    cut = graph["data", "to", "hidden"].edge_index.cpu().numpy()
    got = R.canonical_sort(cut[:, np.isin(cut[1], ts)])
    print("cutoff sample:", "OK" if np.array_equal(got, want) else "MISMATCH", f"({ts.size} targets, {got.shape[1]} edges, radius {radius:.6f})")


if __name__ == "__main__":
    main()
