#!/bin/bash
# one GPU-box session: tests, smoke, bench, ncu launch list (+ optional full captures)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
if [ -n "$NCU_FULL" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$NCU_FULL" -c ${NCU_COUNT:-24} -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
tail -3 gpurun_out/bench_ncu_full.log
fi
