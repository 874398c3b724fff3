#!/bin/bash
# round 2, post-sort tail at N = 2 (sharded output): bit-identity with the gated tail, then A/B of the bench
mkdir -p gpurun_out
TR() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$1"; }
timeout -k 10 200 $(TR 1) tools/sharded_check.py o1280_res7 2>gpurun_out/u_check.err | grep "^{" | cut -c1-400
B() { name=$1; port=$2; shift; shift
  env "$@" timeout -k 10 150 $(TR $port) bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/u_$name.err | grep "^{" > gpurun_out/u_$name.json
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/u_{sys.argv[1]}.json").read())
    print(sys.argv[1], "ms_per_step", d["ms_per_step"], d.get("ms_per_step_by_rank"), "e2e", d["e2e"]["ms_per_step"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
B base 2 AGX_PRELAUNCH_TAIL=0
B new  3 AGX_PRELAUNCH_TAIL=auto
AGX_MIN_SHM_FREE_BYTES=1e30 timeout -k 10 150 $(TR 4) bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/u_noshm.err | grep "^{" | cut -c1-300
ls /dev/shm | head
