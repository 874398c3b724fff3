"""Where does the HOST time of one bench step go?  (run on the GPU box)"""
import cProfile, pstats, io, sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
import bench
from anemoi_graphs_b200 import device as agx_device
from anemoi_graphs_b200.create import GraphCreator

workload = sys.argv[1] if len(sys.argv) > 1 else "o1280_res7"
grid, res = bench.WORKLOADS[workload]
x_host = bench.data_coordinates(grid).pin_memory()
x_dev = x_host.cuda()
creator = GraphCreator(bench.recipe(res))
for resident in (True, False):
    agx_device.set_resident(resident)
    x = x_dev if resident else x_host
    for _ in range(3):
        bench.run_step(creator, x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        bench.run_step(creator, x)
    torch.cuda.synchronize()
    print(f"resident={resident}: {(time.perf_counter()-t0)/5*1e3:.2f} ms/step")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        bench.run_step(creator, x)
    torch.cuda.synchronize()
    pr.disable()
    for key in ("tottime", "cumulative"):
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
        print(s.getvalue()[:6000])
