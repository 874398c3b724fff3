#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N"
echo "== dist_check (gathered mode, forced sharding), peer-mapped exchange (default)"; timeout 300 $TR tools/dist_check.py 2>&1 | grep "dist_check\|Error" | head -5
echo "== dist_check AGX_VMM_PUSH=0 (NCCL)"; AGX_VMM_PUSH=0 timeout 300 $TR tools/dist_check.py 2>&1 | grep "dist_check\|Error" | head -5
for v in 1 0; do
echo "== sweep 100 M, AGX_VMM_PUSH=$v"; AGX_VMM_PUSH=$v timeout 600 $TR tools/scale_sweep.py --queries 100000000 --ks 3,16 --degrees 8,64 2>/dev/null | grep "^{" | tee gpurun_out/vmm_n${N}_push$v.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], d.get('k', d.get('target_degree')), 'sharded', d['ms_sharded'], 'gathered', d.get('ms_gathered'))"
done
