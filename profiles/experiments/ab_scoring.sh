#!/bin/bash
# ab_scoring.sh NAME...: c4 scoring (all 1.6 M users x 54 K items, top-20) for every library variant
for v in "$@"; do
  LGC_B200_LIB=$PWD/gnn_ecommerce_b200/variants/liblgc_$v.so timeout 300 python bench.py --only-scoring 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])['scoring']
print('$v', 'ms', round(d['ms'], 3), {k: round(v, 3) for k, v in d['class_ms'].items()}, 'groups', d.get('candidate_groups'), d.get('check'))"
done
