#!/bin/bash
# ab_variants.sh NAME...: fused-step bench (c2, no CPU leg, no scoring) for every library variant
for v in "$@"; do
  LGC_B200_LIB=$PWD/gnn_ecommerce_b200/variants/liblgc_$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-scoring --no-epoch 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'ms_per_step', round(d['ms_per_step'], 4), {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d['roofline']['class_ms_per_step'].items()}, 'frac', round(d['roofline']['step_frac'], 4))"
done
