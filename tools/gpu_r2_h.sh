#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_o1280_every_knn_edge_vs_sklearn_itself > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/h_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:k_edge_attrs|k_attr_scale|k_node_tables" -c 80 --csv --log-file gpurun_out/h_attr_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/h_ncu.log 2>&1
tail -4 gpurun_out/h_tests.log
python - <<'PY'
import json
for f in ('h_bench',):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); r=d['roofline']; print(f, d['ms_per_step'], d['e2e']['ms_per_step'], {k:v for k,v in r['stage_ms_per_step'].items() if 'attr' in k or 'node' in k})
    except Exception as e: print(f, 'failed', e)
PY
