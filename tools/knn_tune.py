"""KNN decoder kernel on O1280 -> res 7: time + telemetry for a few first-cap sizes (run on the GPU box)."""
import os, sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
import bench
from anemoi_graphs_b200 import ops
from oracle import ref_path as R

x = bench.data_coordinates("o1280").cuda()
hx = torch.from_numpy(R.tri_nodes(7)[0]).cuda()
for cells in (0, 96, 160, 200):
  for scale in ("9", "6", "4.5", "3.5"):
    os.environ["AGX_KNN_CAP_SCALE"] = scale
    with ops.NeighbourIndex(hx, cells_per_face=cells, hint_k=3) as ix:
        out = torch.empty((2, x.shape[0] * 3), dtype=torch.int32, device="cuda")
        for _ in range(2):
            ix.knn(x, 3, out=out)
        stats = ops.new_stats("cuda")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            ix.knn(x, 3, out=out)
        b.record()
        torch.cuda.synchronize()
        ix.knn(x, 3, out=out, stats=stats)
        st = stats.cpu().tolist()
        tiles = (x.shape[0] + 31) // 32
        print(f"cells={ix.cells_per_face} cap_scale={scale}: {a.elapsed_time(b)/5*1e3:.1f} us  f64={st[0]} tie={st[1]} widened={st[2]} staged/tile={st[3]/32/tiles:.1f}", flush=True)
