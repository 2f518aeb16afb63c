#!/bin/bash
# what bounds k_spmm_light: 0 = normal, 1 = no gather loads (zeros), 2 = all gathers hit 16 KB (L1)
for m in 0 5 8 24; do
LGC_LIGHT_DEBUG=$m timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-scoring 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('debug mode $m', 'ms_per_step', round(d['ms_per_step'], 4), 'light', round(d['roofline']['class_ms_per_step']['light'], 4))"
done
