"""Device -> host copy timeline of ONE host-resident build of the headline graph (run on the GPU box): for every copy
when its data was ready on the device, when the copy engine started and finished it, and the idle gaps of the engine.

    python tools/copy_timeline.py
"""
import json
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch

import bench
from anemoi_graphs_b200 import device as agx_device
from anemoi_graphs_b200.create import GraphCreator

grid, res = bench.WORKLOADS["o1280_res7"]
x_host = bench.data_coordinates(grid).pin_memory()
creator = GraphCreator(bench.recipe(res))
agx_device.set_resident(False)
g = None
for _ in range(6):
    g = None
    g = bench.run_step(creator, x_host)
g = None
torch.cuda.synchronize()
for rep in range(2):
    agx_device.copy_trace = []
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record()
    w0 = time.perf_counter()
    g = bench.run_step(creator, x_host)
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - w0)
    rows = []
    for nbytes, ready, begin, done in agx_device.copy_trace:
        rows.append((t0.elapsed_time(ready), t0.elapsed_time(begin), t0.elapsed_time(done), nbytes))
    agx_device.copy_trace = None
    g = None
    rows.sort(key=lambda r: r[1])
    busy, prev_end, gaps = 0.0, None, []
    out = []
    for ready, begin, done, nbytes in rows:
        busy += done - begin
        if prev_end is not None and begin - prev_end > 0.02:
            gaps.append((round(prev_end, 2), round(begin, 2)))
        prev_end = done if prev_end is None else max(prev_end, done)
        out.append({"MB": round(nbytes / 1e6, 1), "ready": round(ready, 2), "begin": round(begin, 2), "done": round(done, 2),
                    "GBps": round(nbytes / 1e6 / max(done - begin, 1e-6), 1)})
    print(json.dumps({"rep": rep, "wall_ms": round(wall, 2), "copies": len(rows), "MB": round(sum(r[3] for r in rows) / 1e6, 1),
                      "engine_busy_ms": round(busy, 2), "first_begin": round(rows[0][1], 2), "last_done": round(max(r[2] for r in rows), 2),
                      "idle_gaps_ms": gaps, "trace": agx_device.last_trace and {k: round(1e3 * (v - w0), 2) for k, v in agx_device.last_trace.items()}}))
    if rep == 1:
        for o in out:
            print(json.dumps(o))
