#!/bin/bash
# round 2, call F: k_knn2 - tests, A/B against the first-generation kernel, ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_o1280_every_knn_edge_vs_sklearn_itself > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
AGX_KNN_V1=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_v1.json 2>> gpurun_out/f_bench.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_knn2 -c 1 -o gpurun_out/f_knn2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu.log 2>&1
tail -4 gpurun_out/f_tests.log
python - <<'PY'
import json
for f in ('f_bench','f_bench_v1'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); r=d['roofline']; print(f, d['ms_per_step'], d['e2e']['ms_per_step'], r['ms_per_launch'], r['frac'], r['fp32'], r['stage_ms_per_step'])
    except Exception as e: print(f, 'failed', e)
PY
ls -la gpurun_out/f_knn2.ncu-rep
