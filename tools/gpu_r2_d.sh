#!/bin/bash
# round 2, call D (2 GPUs): sharded output mode - parity against the single-GPU build, bench at N = 2 and N = 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/sharded_check.py o96_res5 > gpurun_out/d_check_o96.log 2>&1; echo "rc=$?" >> gpurun_out/d_check_o96.log
timeout 600 $TR tools/sharded_check.py o1280_res7 > gpurun_out/d_check_o1280.log 2>&1; echo "rc=$?" >> gpurun_out/d_check_o1280.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/d_bench_n2.json 2> gpurun_out/d_bench_n2.err; echo "rc=$?" >> gpurun_out/d_bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench_n1.json 2> gpurun_out/d_bench_n1.err
tail -4 gpurun_out/d_check_o96.log; tail -4 gpurun_out/d_check_o1280.log; tail -5 gpurun_out/d_bench_n2.err
python - <<'PY'
import json
for f in ('d_bench_n1','d_bench_n2'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['fp32']['frac'], d['roofline']['stage_ms_per_step'])
    except Exception as e: print(f, 'failed', e)
PY
