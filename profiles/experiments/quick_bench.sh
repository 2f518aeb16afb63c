#!/bin/bash
# parity of the SpMM paths, then the fused-step bench without the CPU / scoring legs
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-scoring 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'], 4), d['roofline']['class_ms_per_step'], 'frac', round(d['roofline']['step_frac'], 4))"
