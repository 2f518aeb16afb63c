#!/bin/bash
# r2h: ncu --set full of the scoring kernels of the shipped build (one c4 call: 1.6 M users x 54 K items, top-20)
CMD="python bench.py --only-scoring"
$CMD > gpurun_out/plain_scoring.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_score_gemm|k_threshold|k_scan|k_rescore|k_select|k_exhaustive" -c 7 -f -o gpurun_out/r2h_scoring $CMD > gpurun_out/ncu_scoring.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_scoring.log
