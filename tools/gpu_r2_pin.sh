#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
for v in "" "AGX_PIN_SORT=1" "AGX_SHARED_ORDER=0"; do
  echo "== N=2 $v"
  env $v timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
done
echo "== N=1"; timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
echo "== N=1 AGX_PIN_SORT=1"; AGX_PIN_SORT=1 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
nproc; lscpu | grep -i "thread\|core\|socket\|model name"
