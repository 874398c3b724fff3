"""Host timeline of the provisional node order inside one bench step (run on the GPU box; under torchrun every rank
prints its own line: sharded output mode)."""
import os, sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
import bench
from anemoi_graphs_b200 import device as agx_device
from anemoi_graphs_b200.create import GraphCreator

world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
if world > 1:
    import torch.distributed as dist

    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    agx_device.set_sharded_output(os.environ.get('AGX_TRACE_SHARDED', '1') == '1')
grid, res = bench.WORKLOADS["o1280_res7"]
x_host = bench.data_coordinates(grid).pin_memory()
x_dev = x_host.cuda()
creator = GraphCreator(bench.recipe(res))
for resident in (True,) if os.environ.get('AGX_TRACE_RESIDENT_ONLY') else (True, False):
    agx_device.set_resident(resident)
    x = x_dev if resident else x_host
    back_to_back = os.environ.get("AGX_TRACE_BACK_TO_BACK") == "1"  # as bench.py runs them: no sync between steps
    for rep in range(8):
        g = None
        if not back_to_back or rep == 0:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = bench.run_step(creator, x)
        t1 = time.perf_counter()
        if not back_to_back:
            torch.cuda.synchronize()
        t2 = time.perf_counter()
        tr = agx_device.last_trace
        if rep >= 5:
            print(f"rank {rank} resident={resident} step: returned {1e3*(t1-t0):.2f} ms, synced {1e3*(t2-t0):.2f} ms | " +
                  "  ".join(f"{k}={1e3*(v-t0):.2f}" for k, v in tr.items()), flush=True)
if world > 1:
    dist.destroy_process_group()
