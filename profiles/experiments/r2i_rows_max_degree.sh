#!/bin/bash
# r2i: degree above which a row goes to the sweep instead of the rows kernel (LGC_ROWS_MAX_DEGREE, default 16), c2 step
for d in ${DEGREES:-8 16 24 32 48}; do
  LGC_ROWS_MAX_DEGREE=$d timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-scoring --no-epoch 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
c = d['roofline']['class_ms_per_step']
print('max_degree=$d', 'ms_per_step', round(d['ms_per_step'], 4), 'rows', round(c['light'], 4), 'sweep', round(c['heavy'], 4), 'finish', round(c['finish'], 4), 'losses', d['losses_last_step'][0])"
done
