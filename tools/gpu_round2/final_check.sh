#!/bin/bash
# round 2, last call: the builder-API GPU tests, smoke() and the default bench line with the final defaults
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_builders.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 10 300 python bench.py --steps 20 --warmup 5 2> gpurun_out/final_bench.err | grep "^{" > gpurun_out/final_bench.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_bench.json").read())
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["fp32"]["frac"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"]["value"])
PY
