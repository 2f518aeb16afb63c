#!/bin/bash
# A/B of the light-row kernel variants inside one box: parity first, then the fused-step bench.
set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for b in 0 1; do
  LGC_LIGHT_BALANCED=$b timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-scoring 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('balanced=$b', 'ms_per_step', d['ms_per_step'], d['roofline']['class_ms_per_step'], 'frac', d['roofline']['step_frac'])"
done
