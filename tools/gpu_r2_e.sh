#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
export AGX_TRACE_RESIDENT_ONLY=1
echo "== N=1" > gpurun_out/e_trace.log; timeout 300 python tools/lazy_trace.py 2>&1 | grep "^rank" >> gpurun_out/e_trace.log
echo "== N=1 with OMP_NUM_THREADS=1" >> gpurun_out/e_trace.log; OMP_NUM_THREADS=1 timeout 300 python tools/lazy_trace.py 2>&1 | grep "^rank" >> gpurun_out/e_trace.log
echo "== N=2 sharded, shared order" >> gpurun_out/e_trace.log; timeout 300 $TR tools/lazy_trace.py 2>&1 | grep "^rank" >> gpurun_out/e_trace.log
echo "== N=2 sharded, own sorts" >> gpurun_out/e_trace.log; AGX_SHARED_ORDER=0 timeout 300 $TR tools/lazy_trace.py 2>&1 | grep "^rank" >> gpurun_out/e_trace.log
echo "== N=2 replicated builds (gather mode: below the shard threshold nothing is exchanged), own sorts" >> gpurun_out/e_trace.log; AGX_SHARED_ORDER=0 AGX_TRACE_SHARDED=0 timeout 300 $TR tools/lazy_trace.py 2>&1 | grep "^rank" >> gpurun_out/e_trace.log
echo "== two independent single-GPU processes side by side" >> gpurun_out/e_trace.log
(CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/lazy_trace.py 2>&1 | grep "^rank" > gpurun_out/e_trace_a.log) &
(CUDA_VISIBLE_DEVICES=1 timeout 300 python tools/lazy_trace.py 2>&1 | grep "^rank" > gpurun_out/e_trace_b.log) &
wait
cat gpurun_out/e_trace_a.log gpurun_out/e_trace_b.log >> gpurun_out/e_trace.log
lscpu | grep -i "model name\|thread\|core\|socket\|numa" >> gpurun_out/e_trace.log
cat gpurun_out/e_trace.log | cut -c1-330
