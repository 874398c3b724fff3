#!/bin/bash
# quick A/B on the GPU box: tests, then bench under a few tuning knobs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for j in 2; do
  echo "== AGX_ATTR_J=$j"
  AGX_ATTR_J=$j timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['stage_ms_per_step'])"
done
