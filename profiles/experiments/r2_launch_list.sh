#!/bin/bash
# every launch of one fused c2 step with its device time (ncu serialises and runs cold: compare shares)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scoring"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 40 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]:
    n=r[ki]; n=n[n.find('k_'):][:44] if 'k_' in n else n[:44]
    print(f"{n:46s} {r[vi]}")
PY
