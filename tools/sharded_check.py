"""Multi-GPU parity of the SHARDED output mode (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_check.py [workload]

Every rank builds the workload's graph (a) alone, as a single-GPU build (``device.single_rank()``), (b) sharded with the
result on the HOST - one complete graph in the node-wide shared page-locked buffer, each rank having written its block
over its own PCIe link - and (c) sharded and device-resident (each rank keeps its block).  Checked on every rank:
(b) equals (a) column for column (edge lists bit-exact, attributes to 3e-7: the normalisation statistics are folded per
rank), the blocks of (c) are exactly the matching columns of (a), and all ranks see the same host bytes.
Prints one JSON line on rank 0; exit code 1 on any mismatch."""

import json
import os
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench
from anemoi_graphs_b200 import device as D
from anemoi_graphs_b200.create import GraphCreator
from anemoi_graphs_b200.graph import HeteroData

workload = sys.argv[1] if len(sys.argv) > 1 else "o96_res5"
grid, res = bench.WORKLOADS[workload]
x_host = bench.data_coordinates(grid).pin_memory()
x_dev = x_host.cuda()
creator = GraphCreator(bench.recipe(res))
NAMES = ("edge_index", "edge_length", "edge_dirs")


def build(x):
    g = HeteroData()
    g["data"].x = x
    g["data"].node_type = "LatLonNodes"
    return creator.update_graph(g)


problems = []


def check(cond, what):
    if not cond:
        problems.append(f"rank {rank}: {what}")


# (a) single-GPU build on every rank
D.set_sharded_output(False)
with D.single_rank():
    D.set_resident(False)
    single = build(x_host)
    torch.cuda.synchronize()

# (b) sharded, host-resident
D.set_sharded_output(True)
D.set_resident(False)
for rep in range(3):  # repeated: the shared segments are reused from build to build
    sharded = None
    sharded = build(x_host)
    torch.cuda.synchronize()
check(np.array_equal(sharded["hidden"].x.numpy().view(np.int32), single["hidden"].x.numpy().view(np.int32)), "hidden x")
check(np.array_equal(np.asarray(sharded["hidden"]["_node_ordering"]), np.asarray(single["hidden"]["_node_ordering"])), "node ordering")
for key in bench.EDGE_KEYS:
    a, b = single[key], sharded[key]
    check(tuple(a.edge_index.shape) == tuple(b.edge_index.shape), f"{key} edge count {tuple(b.edge_index.shape)}")
    check(torch.equal(a.edge_index, b.edge_index), f"{key} edge_index differs from the single-GPU build")
    for name in NAMES[1:]:
        va, vb = a[name].numpy(), b[name].numpy()
        check(va.shape == vb.shape and np.allclose(vb, va, rtol=3e-7, atol=1e-7 * np.abs(va).max()), f"{key} {name}")
    # every rank sees the same host bytes (one shared copy)
    for name in NAMES:
        t = b[name].contiguous().view(torch.uint8).to(torch.int64).sum().cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        check(lo.item() == hi.item(), f"{key} {name}: ranks see different host bytes")

# (c) sharded, device-resident: this rank's blocks are the matching columns of the single-GPU result
D.set_resident(True)
blocks = build(x_dev)
torch.cuda.synchronize()
for key in bench.EDGE_KEYS:
    info = blocks[key]["edge_shard"]
    check(info["rank"] == rank and info["world"] == world, f"{key} shard info {info}")
    a = single[key]
    if info["replicated"]:
        lo, hi = 0, int(a.edge_index.shape[1])
    else:
        lo = sum(info["counts"][:rank])
        hi = lo + info["counts"][rank]
        check(sum(info["counts"]) == int(a.edge_index.shape[1]), f"{key} counts {info['counts']}")
    check(torch.equal(blocks[key].edge_index.cpu(), a.edge_index[:, lo:hi]), f"{key} resident block [{lo}:{hi})")
    for name in NAMES[1:]:
        va, vb = a[name][lo:hi].numpy(), blocks[key][name].cpu().numpy()
        check(va.shape == vb.shape and np.allclose(vb, va, rtol=3e-7, atol=1e-7 * np.abs(a[name].numpy()).max()), f"{key} {name} resident block")
D.set_resident(False)

# quick timing of the two sharded modes (bench.py is the measurement; this is a smoke number)
def timed(fn, n=5):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        g = None
        g = fn()
    torch.cuda.synchronize()
    dist.barrier()
    return (time.perf_counter() - t0) / n * 1e3


from anemoi_graphs_b200 import shm

arena_before = shm.arena().stats()
ms_host_first = timed(lambda: build(x_host), n=1)
arena_first = shm.arena().stats()
ms_host = timed(lambda: build(x_host))
arena_after = shm.arena().stats()
D.set_resident(True)
ms_res = timed(lambda: build(x_dev))
D.set_resident(False)

flag = torch.tensor([0 if problems else 1], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
every = [None] * world
dist.all_gather_object(every, problems)
if rank == 0:
    print(json.dumps({"tool": "sharded_check", "workload": workload, "world": world, "ok": bool(flag.item()),
                      "edges": {str(k): int(single[k].edge_index.shape[1]) for k in bench.EDGE_KEYS},
                      "ms_per_build_host_out": round(ms_host, 3), "ms_first_host_out": round(ms_host_first, 3),
                      "ms_per_build_resident": round(ms_res, 3), "arena": [arena_before, arena_first, arena_after],
                      "problems": [p for ps in every for p in ps][:20]}))  # fmt: skip
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
