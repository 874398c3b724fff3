#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full capture of the search / attribute kernels,
# summarised ON the box (the .ncu-rep with source is too large to travel back).
#   TAG=r01_v4 bash tools/gpu_profile.sh
set -x
TAG=${TAG:-rXX}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches.csv gpurun_out/${TAG}_launches.txt > /dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${NCU_FULL:-k_knn|k_radius|k_edge_attrs|k_attr_scale}" \
    -c ${NCU_COUNT:-32} -f -o /tmp/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
python tools/ncu_summary.py full /tmp/prof.ncu-rep gpurun_out/${TAG}_ncu_full_summary.csv > /dev/null
ls -la /tmp/prof.ncu-rep gpurun_out
head -30 gpurun_out/${TAG}_launches.txt
