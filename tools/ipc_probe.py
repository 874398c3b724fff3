"""Probe: how fast can one rank write into another rank's buffer through PyTorch's CUDA IPC tensor sharing?
(run under torchrun with 2 ranks on the GPU box)"""
import os, sys, time
import torch, torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 * 1024 * 1024  # 1 GiB of int32
buf = torch.zeros(n, dtype=torch.int32, device="cuda")
src = torch.full((n,), rank + 1, dtype=torch.int32, device="cuda")
t0 = time.perf_counter()
fn, args = reduce_tensor(buf)
gathered = [None] * world
dist.all_gather_object(gathered, args)
t1 = time.perf_counter()
peer = (rank + 1) % world
variants = {}
# (a) as PyTorch rebuilds it: a tensor on the producer's device index
variants["peer-device tensor"] = fn(*gathered[peer])
# (b) the same handle opened in MY device's context: a tensor that claims to live on my device
a = list(gathered[peer])
a[6] = local  # storage_device
try:
    variants["opened on my device"] = fn(*a)
except Exception as e:  # noqa
    print(rank, "variant b failed:", repr(e)[:200])
torch.cuda.synchronize()
t2 = time.perf_counter()
if rank == 0:
    print(f"share+all_gather_object {1e3*(t1-t0):.2f} ms, open {1e3*(t2-t1):.2f} ms")
for name, view in variants.items():
    for rep in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        view.copy_(src, non_blocking=True)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        if rank == 0 and rep > 0:
            print(f"{name}: device {view.device}, 1 GiB push {ms:.2f} ms = {4*n/ms/1e6:.0f} GB/s")
    dist.barrier()
    torch.cuda.synchronize()
    ok = bool((buf == ((rank - 1) % world) + 1).all().item())
    if rank == 0:
        print(f"{name}: received correctly: {ok}")
    buf.zero_()
    torch.cuda.synchronize()
    dist.barrier()
# NCCL reference: all_gather of the same bytes
out = torch.empty(world * n, dtype=torch.int32, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); dist.all_gather_into_tensor(out, src); e.record(); torch.cuda.synchronize()
    if rank == 0 and rep > 0:
        print(f"nccl all_gather 1 GiB/rank: {s.elapsed_time(e):.2f} ms")
del variants
dist.destroy_process_group()
