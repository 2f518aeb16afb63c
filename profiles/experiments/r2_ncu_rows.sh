#!/bin/bash
# ncu --set full of the rows kernel: the 6 launches of one fused step after warm-up
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scoring"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_spmm_rows -s 18 -c 6 -f -o gpurun_out/r2_rows $CMD > gpurun_out/ncu_rows.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_rows.log
