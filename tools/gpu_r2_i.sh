#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_config4.py tests/test_gpu_kernels.py tests/test_gpu_builders.py -m gpu -x -q --durations=12 > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/i_tests.log
tail -40 gpurun_out/i_tests.log
