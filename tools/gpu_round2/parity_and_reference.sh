#!/bin/bash
# round 2, call A: parity at full size, boundary enumeration, baseline bench, the reference's full pass
mkdir -p gpurun_out
nproc > gpurun_out/a_host.txt; free -g >> gpurun_out/a_host.txt; nvidia-smi -L >> gpurun_out/a_host.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
timeout 600 python tools/boundary_cases.py --out gpurun_out/o1280_boundary_cases.json > gpurun_out/a_boundary.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
timeout 900 python bench.py --impl reference > gpurun_out/a_reference.json 2> gpurun_out/a_reference.err
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_cpu.json 2>> gpurun_out/a_bench.err
tail -5 gpurun_out/a_tests.log; cat gpurun_out/a_boundary.log | tail -3; cut -c1-600 gpurun_out/a_bench.json; cut -c1-400 gpurun_out/a_reference.json
