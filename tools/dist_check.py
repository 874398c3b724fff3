"""Multi-GPU parity: under torchrun, every rank builds the O96 -> res 5 graph (config 1) with sharded queries /
edges and must end with EXACTLY the single-GPU result (golden fixtures of the unmodified reference)."""
import os, sys, pathlib
os.environ.setdefault("AGX_SHARD_MIN_QUERIES", "0")  # force the sharded paths: this graph is far below the thresholds
os.environ.setdefault("AGX_ATTR_SHARD_MIN_EDGES", "0")
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from anemoi_graphs_b200 import grids
from anemoi_graphs_b200.create import GraphCreator
from anemoi_graphs_b200.graph import HeteroData
import bench

g = np.load(pathlib.Path(__file__).resolve().parents[1] / "tests" / "golden" / "o96_res5.npz")
lat, lon = grids.octahedral_grid(96)
graph = HeteroData()
graph["data"].x = grids.latlon_deg_to_x(lat, lon)
graph["data"].node_type = "LatLonNodes"
graph = GraphCreator(bench.recipe(5)).update_graph(graph)


def canon(ei):
    ei = ei.numpy()
    return ei[:, np.lexsort((ei[0], ei[1]))]


ok = True
cut, ms, knn = (canon(graph[k].edge_index) for k in bench.EDGE_KEYS)
ok &= np.array_equal(cut, g["cutoff_edge_index"]) and np.array_equal(ms, g["multiscale_edge_index"])
ok &= knn.shape == g["knn3_edge_index"].shape
stride = int(g["attr_sample_stride"])
for key, tag in zip(bench.EDGE_KEYS[:2], ("cutoff", "multiscale")):
    ei = graph[key].edge_index.numpy()
    order = np.lexsort((ei[0], ei[1]))
    ln = graph[key]["edge_length"].numpy()[order][::stride]
    dr = graph[key]["edge_dirs"].numpy()[order][::stride]
    ok &= np.allclose(ln, g[f"{tag}_edge_length_sample"], rtol=1e-6, atol=0)
    want = g[f"{tag}_edge_dirs_sample"]
    ok &= np.allclose(dr, want, rtol=1e-6, atol=1e-6 * np.abs(want).max())
# every rank must hold bit-identical tensors
for key in bench.EDGE_KEYS:
    for name in ("edge_index", "edge_length", "edge_dirs"):
        t = graph[key][name].cuda().contiguous().view(torch.uint8).to(torch.int64).sum()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok &= bool(lo.item() == hi.item())
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"dist_check world={world}: {'OK' if flag.item() else 'MISMATCH'} edges={cut.shape[1]},{ms.shape[1]},{knn.shape[1]}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
