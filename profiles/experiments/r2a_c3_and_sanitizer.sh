#!/bin/bash
# r2a: c3 (K=5, d=90) on one GPU -- never measured in round 1 -- and compute-sanitizer over the c1-sized GPU tests
mkdir -p gpurun_out
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline --no-scoring \
  > gpurun_out/r2a_bench_c3_n1.json 2> gpurun_out/r2a_bench_c3_n1.log
echo "c3 bench rc=$?"; cat gpurun_out/r2a_bench_c3_n1.json | head -c 3000
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "golden_tiny or golden_c1 or edge_cases or hub_rows" > gpurun_out/r2a_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r2a_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "fused_step_matches_golden_tiny or hub_rows" > gpurun_out/r2a_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/r2a_sanitizer_racecheck.log
