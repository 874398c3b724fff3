"""Wall-clock timeline of one bench step, stage by stage (synchronising after each stage) - run on the GPU box."""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
import bench
from anemoi_graphs_b200 import device as agx_device
from anemoi_graphs_b200.config import DotDict, instantiate
from anemoi_graphs_b200.graph import HeteroData

import os
if int(os.environ.get("WORLD_SIZE", "1")) > 1:  # under torchrun: sharded build, rank 0 reports
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    if int(os.environ["RANK"]) != 0:
        sys.stdout = open(os.devnull, "w")
workload = sys.argv[1] if len(sys.argv) > 1 else "o1280_res7"
grid, res = bench.WORKLOADS[workload]
x_host = bench.data_coordinates(grid).pin_memory()
x_dev = x_host.cuda()
cfg = DotDict(bench.recipe(res))
for resident in (True, False):
    agx_device.set_resident(resident)
    x = x_dev if resident else x_host
    acc = {}
    n_rep = 8
    for rep in range(n_rep + 2):
        graph = HeteroData()
        graph["data"].x = x
        graph["data"].node_type = "LatLonNodes"
        torch.cuda.synchronize()
        marks = []
        t0 = time.perf_counter()

        def mark(name):
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

        with agx_device.deferred():
            for nodes_name, nodes_cfg in cfg.nodes.items():
                graph = instantiate(nodes_cfg.node_builder, name=nodes_name).update_graph(graph, attrs_config={})
                mark(f"nodes:{nodes_name}")
            for edges_cfg in cfg.edges:
                for b in edges_cfg.edge_builders:
                    eb = instantiate(b, source_name=edges_cfg.source_name, target_name=edges_cfg.target_name)
                    graph = eb.update_graph(graph, attrs_config=None)
                    mark(f"edges:{edges_cfg.source_name}->{edges_cfg.target_name}")
                graph = eb.register_attributes(graph, edges_cfg.get("attributes", {}))
                mark(f"attrs:{edges_cfg.source_name}->{edges_cfg.target_name}")
        mark("flush")
        if rep >= 2:
            prev = t0
            for name, t in marks:
                acc[name] = acc.get(name, 0.0) + (t - prev)
                prev = t
    print(f"resident={resident}: " + "  ".join(f"{k}={v/n_rep*1e3:.2f}ms" for k, v in acc.items()) + f"  total={sum(acc.values())/n_rep*1e3:.2f}ms")
