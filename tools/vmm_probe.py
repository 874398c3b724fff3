"""Probe for the round-2 exchange: can one rank write into another rank's buffer at NVLink
speed through the driver's virtual-memory API (cuMemCreate + a POSIX-fd shareable handle passed over a Unix socket +
cuMemMap / cuMemSetAccess for the LOCAL device)?  `tools/ipc_probe.py` showed that PyTorch's legacy CUDA-IPC route
moves only 26 GB/s on these boxes.  Run with 2 ranks:

    python -m torch.distributed.run --nproc-per-node 2 tools/vmm_probe.py
"""
import os
import socket
import sys
import time

import torch
import torch.distributed as dist
from cuda.bindings import driver as cu

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
assert world == 2, "two ranks"
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.zeros(1, device="cuda")  # primary context


def ck(res):
    err = res[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"rank {rank}: {err}")
    return res[1] if len(res) == 2 else res[1:]


SIZE = 1 << 30
prop = cu.CUmemAllocationProp()
prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
prop.location.id = local
prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
gran = ck(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
size = (SIZE + gran - 1) // gran * gran


def map_here(handle):
    va = ck(cu.cuMemAddressReserve(size, 0, 0, 0))
    ck(cu.cuMemMap(va, size, 0, handle, 0) + (None,))
    acc = cu.CUmemAccessDesc()
    acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = local
    acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    ck(cu.cuMemSetAccess(va, size, [acc], 1) + (None,))
    return va


own = ck(cu.cuMemCreate(size, prop, 0))
own_va = map_here(own)
fd = ck(cu.cuMemExportToShareableHandle(own, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0))
fd = int(fd)

# swap the file descriptors over a Unix socket (SCM_RIGHTS)
path = f"/tmp/agx_vmm_probe_{os.environ.get('MASTER_PORT', '0')}.sock"
if rank == 0:
    if os.path.exists(path):
        os.unlink(path)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    srv.listen(1)
    dist.barrier()
    conn, _ = srv.accept()
else:
    dist.barrier()
    conn = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    for _ in range(100):
        try:
            conn.connect(path)
            break
        except OSError:
            time.sleep(0.05)
socket.send_fds(conn, [b"x"], [fd])
_, fds, _, _ = socket.recv_fds(conn, 16, 1)
peer = ck(cu.cuMemImportFromShareableHandle(fds[0], cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR))
t0 = time.perf_counter()
peer_va = map_here(peer)  # the PEER's physical memory, mapped for access from MY device
t_map = time.perf_counter() - t0

src = torch.full((size // 4,), rank + 1, dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
ck(cu.cuMemsetD32Async(own_va, 0, size // 4, stream) + (None,))
torch.cuda.synchronize()
dist.barrier()
for rep in range(4):
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ck(cu.cuMemcpyDtoDAsync(peer_va, src.data_ptr(), size, stream) + (None,))
    b.record()
    torch.cuda.synchronize()
    if rank == 0 and rep > 0:
        ms = a.elapsed_time(b)
        print(f"push 1 GiB into the peer's VMM mapping: {ms:.2f} ms = {size / ms / 1e6:.0f} GB/s (both ranks pushing at once)")
dist.barrier()
torch.cuda.synchronize()
# what arrived in MY buffer must be the peer's value
mine = torch.empty(size // 4, dtype=torch.int32, device="cuda")
ck(cu.cuMemcpyDtoDAsync(mine.data_ptr(), own_va, size, stream) + (None,))
torch.cuda.synchronize()
ok = bool((mine == (1 - rank) + 1).all().item())
if rank == 0:
    print(f"mapping the peer allocation took {1e3 * t_map:.2f} ms; received correctly: {ok}")
dist.barrier()
dist.destroy_process_group()
