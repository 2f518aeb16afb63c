#!/bin/bash
# ncu --set full of the sweep kernel (3 launches after warm-up) in a short fused-step run
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scoring"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_spmm_sweep -s 18 -c 3 -f -o gpurun_out/r2_sweep $CMD > gpurun_out/ncu_sweep.log 2>&1
echo rc=$?; tail -3 gpurun_out/ncu_sweep.log
