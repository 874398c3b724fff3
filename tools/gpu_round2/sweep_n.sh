#!/bin/bash
# usage: bash tools/gpu_round2/sweep_n.sh N
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N"
timeout 1500 $TR tools/scale_sweep.py 2> gpurun_out/sweep_n$N.err | grep "^{" > gpurun_out/sweep_n$N.jsonl
if [ "$N" = "2" ]; then
  AGX_TRACE_RESIDENT_ONLY=1 AGX_TRACE_BACK_TO_BACK=1 timeout 300 $TR tools/lazy_trace.py 2>&1 | grep "^rank" > gpurun_out/trace_b2b_n2.log
  AGX_TRACE_RESIDENT_ONLY=1 AGX_TRACE_BACK_TO_BACK=1 timeout 300 python tools/lazy_trace.py 2>&1 | grep "^rank" > gpurun_out/trace_b2b_n1.log
  cut -c1-400 gpurun_out/trace_b2b_n1.log | tail -3; cut -c1-420 gpurun_out/trace_b2b_n2.log | tail -6
fi
tail -2 gpurun_out/sweep_n$N.err; wc -l gpurun_out/sweep_n$N.jsonl; tail -4 gpurun_out/sweep_n$N.jsonl | cut -c1-300
