"""Interleaved A/B of the node-order tail variants on one GPU (run on the GPU box): upload hand-over on / off x
pre-launched (gated) tail on / off, in ONE process, modes alternating block by block so that host drift cancels.
Per mode: median latency of a single synchronised build, and median ms/step of back-to-back blocks (how bench.py times).

    python tools/tail_ab.py [blocks] [steps_per_block]
"""
import json
import pathlib
import statistics
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch

import bench
from anemoi_graphs_b200 import device as agx_device
from anemoi_graphs_b200.create import GraphCreator

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
grid, res = bench.WORKLOADS["o1280_res7"]
x_host = bench.data_coordinates(grid).pin_memory()
x_dev = x_host.cuda()
creator = GraphCreator(bench.recipe(res))
MODES = [(0, 0), (1, 0), (0, 1), (1, 1)]  # (hand-over, pre-launched tail)

for resident in (True, False):
    agx_device.set_resident(resident)
    x = x_dev if resident else x_host
    single = {m: [] for m in MODES}
    block_ms = {m: [] for m in MODES}
    for b in range(blocks + 1):  # block 0 warms up every mode
        for m in MODES:
            agx_device.UPLOAD_HANDOVER, agx_device.PRELAUNCH_TAIL = bool(m[0]), bool(m[1])
            g = None
            for _ in range(3):  # single synchronised builds
                g = None
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                g = bench.run_step(creator, x)
                torch.cuda.synchronize()
                if b:
                    single[m].append(1e3 * (time.perf_counter() - t0))
            g = None
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                g = None
                g = bench.run_step(creator, x)
            e.record()
            torch.cuda.synchronize()
            if b:
                block_ms[m].append(a.elapsed_time(e) / steps)
    for m in MODES:
        print(json.dumps({
            "resident": resident, "handover": m[0], "prelaunch": m[1],
            "single_ms_median": round(statistics.median(single[m]), 3), "single_ms_min": round(min(single[m]), 3),
            "back_to_back_ms_median": round(statistics.median(block_ms[m]), 3),
            "back_to_back_ms_all": [round(v, 3) for v in block_ms[m]],
        }), flush=True)
