#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches.csv gpurun_out/r02_launches.txt > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:k_knn|k_radius|k_edge_attrs|k_attr_scale|k_relabel_rows|k_index" -c 48 -f -o /tmp/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
python tools/ncu_summary.py full /tmp/prof.ncu-rep gpurun_out/r02_ncu_full_summary.csv > /dev/null 2>&1
tail -15 gpurun_out/final_tests.log; cat gpurun_out/final_smoke.log | tail -3; head -40 gpurun_out/r02_launches.txt
