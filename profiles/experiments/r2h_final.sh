#!/bin/bash
# r2h_final.sh: the round's closing single-GPU run -- GPU tests, smoke, bench (product + reference arm), then the ncu
# launch list of one fused step and the --set full capture of the two SpMM kernels (each only after the plain run)
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_gputest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2h_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2h_smoke.log
python bench.py > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.log; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2h_ref_n1.json 2> gpurun_out/r2h_ref_n1.log; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scoring --no-epoch"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 88 -c 44 --csv --log-file gpurun_out/r2h_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_spmm_rows|k_spmm_sweep" -s 36 -c 12 -f -o gpurun_out/r2h_step $CMD > gpurun_out/ncu_step.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_step.log
