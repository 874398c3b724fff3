#!/bin/bash
# round 2: the 1 -> 8 curve on ONE 8-GPU box, the sharded-output parity check at N = 8, the host DMA ceiling
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s_topo.txt 2>&1; nproc >> gpurun_out/s_topo.txt
TR() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2951$1"; }
timeout 600 $(TR 8) tools/sharded_check.py o1280_res7 2>gpurun_out/s_check_n8.err | grep "^{" > gpurun_out/s_check_n8.jsonl
timeout 300 $(TR 8) tools/pcie_probe.py 2>/dev/null | grep "^{" > gpurun_out/s_pcie.jsonl
timeout 300 $(TR 2) tools/pcie_probe.py 2>/dev/null | grep "^{" >> gpurun_out/s_pcie.jsonl
timeout 300 python tools/pcie_probe.py 2>/dev/null | grep "^{" >> gpurun_out/s_pcie.jsonl
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/s_bench_n1.err | grep "^{" > gpurun_out/s_bench_n1.json
  else timeout 600 $(TR $n) bench.py --gpus $n --steps 20 --warmup 5 2> gpurun_out/s_bench_n$n.err | grep "^{" > gpurun_out/s_bench_n$n.json; fi
done
cat gpurun_out/s_check_n8.jsonl | cut -c1-330; cat gpurun_out/s_pcie.jsonl
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open('gpurun_out/s_bench_n%d.json'%n).read()); print(n, d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['ms_per_step_by_rank'], d['roofline']['frac'], d['roofline']['fp32']['frac'])
    except Exception as e: print(n, 'failed', e)
PY
