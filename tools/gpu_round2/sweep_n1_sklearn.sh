#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python tools/scale_sweep.py --sklearn > gpurun_out/sweep_n1.jsonl 2> gpurun_out/sweep_n1.err
tail -3 gpurun_out/sweep_n1.err; cut -c1-260 gpurun_out/sweep_n1.jsonl
