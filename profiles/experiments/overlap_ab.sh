#!/bin/bash
# light / heavy kernels sequential (0) vs concurrent on two streams (1): parity, then step time
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for o in 0 1; do
LGC_SPMM_OVERLAP=$o timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-scoring 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('overlap=$o ms_per_step', round(d['ms_per_step'], 4), d['roofline']['class_ms_per_step'], 'frac', round(d['roofline']['step_frac'], 4))"
done
