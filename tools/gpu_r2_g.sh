#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_builders.py -m gpu -x -q > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g_tests.log
timeout 600 python tools/knn_ab.py > gpurun_out/g_knn_ab.log 2>&1
tail -3 gpurun_out/g_tests.log; cat gpurun_out/g_knn_ab.log
