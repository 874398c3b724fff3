#!/bin/bash
# round 2, call C: the shortened post-sort tail
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_o1280_every_knn_edge_vs_sklearn_itself > gpurun_out/c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c_tests.log
timeout 300 python tools/lazy_trace.py > gpurun_out/c_trace.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
tail -5 gpurun_out/c_tests.log; cat gpurun_out/c_trace.log | tail -12; cut -c1-300 gpurun_out/c_bench.json; python -c "
import json
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['stage_ms_per_step'])
"
