#!/bin/bash
# round 2, post-sort tail: A/B of the upload hand-over and of the pre-launched (gated) tail on one GPU
mkdir -p gpurun_out
timeout -k 10 240 python -m pytest tests/test_gpu_builders.py -m gpu -x -q -k "provisional or failed_node_order or resident or toy or o96" 2>&1 | tail -4
B() { # name, env...
  name=$1; shift
  env "$@" timeout -k 10 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/t_$name.err | grep "^{" > gpurun_out/t_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/t_{sys.argv[1]}.json").read())
    print(sys.argv[1], "ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
B base    AGX_UPLOAD_HANDOVER=0 AGX_PRELAUNCH_TAIL=0
B handover AGX_UPLOAD_HANDOVER=1 AGX_PRELAUNCH_TAIL=0
B gated   AGX_UPLOAD_HANDOVER=1 AGX_PRELAUNCH_TAIL=1
B gated2  AGX_UPLOAD_HANDOVER=1 AGX_PRELAUNCH_TAIL=1
AGX_UPLOAD_HANDOVER=1 AGX_PRELAUNCH_TAIL=1 timeout -k 10 120 python tools/lazy_trace.py 2>&1 | grep "^rank" | tail -4
