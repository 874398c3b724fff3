"""Host DMA ceiling of a box (run alone or under torchrun): every rank copies 1 GiB device -> host, all ranks at once,
into (a) its own cudaHostAlloc'ed buffer and (b) its slice of ONE shared page-locked segment (``shm.HostArena``, what
the sharded output mode writes into).  Prints per-rank and aggregate GB/s - the bound of `e2e` at N > 1."""
import json, os, sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 28  # float32 elements = 1 GiB
src = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
own = torch.empty(n, dtype=torch.float32, pin_memory=True)
targets = {"own_pinned": own}
if world > 1:
    from anemoi_graphs_b200 import shm
    shm.local_group()
    shared = shm.arena().tensor((world, n), torch.float32)
    targets["shared_segment"] = shared[rank]


def run(dst):
    times = []
    for rep in range(4):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        mine = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        total = time.perf_counter() - t0
        if rep:
            times.append((mine, total))
    return min(t[0] for t in times), min(t[1] for t in times)


out = {"tool": "pcie_probe", "world": world, "gib_per_rank": 1}
for name, dst in targets.items():
    mine, total = run(dst)
    gbs = 4 * n / mine / 1e9
    if world > 1:
        t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
        every = torch.empty(world, device="cuda", dtype=torch.float64)
        dist.all_gather_into_tensor(every, t)
        out[name] = {"per_rank_gbs": [round(float(v), 1) for v in every.tolist()], "aggregate_gbs": round(world * 4 * n / total / 1e9, 1)}
    else:
        out[name] = {"per_rank_gbs": [round(gbs, 1)], "aggregate_gbs": round(gbs, 1)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
