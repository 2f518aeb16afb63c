#!/bin/bash
# light-row kernel time vs resident CTAs per SM (persistent grid = 148 x k)
for k in 1 2 3 4 5; do
LGC_LIGHT_CTAS=$k timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-scoring 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ctas/SM $k', 'ms_per_step', round(d['ms_per_step'], 4), 'light', round(d['roofline']['class_ms_per_step']['light'], 4))"
done
