#!/bin/bash
# ncu --set full of the two SpMM kernels of the final build: the 6 + 6 launches of one fused c2 step
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scoring --no-epoch"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_spmm_rows|k_spmm_sweep" -s 36 -c 12 -f -o gpurun_out/r2_step $CMD > gpurun_out/ncu_step.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_step.log
