#!/bin/bash
# r2e_peer_exchange.sh N: the 2-rank GPU tests (peer-memory exchange == ncclAllReduce + epilogue, bit for bit) and one
# c2 bench line per exchange kind -> gpurun_out/r2e_*
N=${1:-2}
timeout 420 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r2e_multirank_test.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2e_multirank_test.log
for X in peer nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 --config c2 --no-scoring --exchange $X > gpurun_out/r2e_c2_n${N}_$X.json 2> gpurun_out/r2e_c2_n${N}_$X.log
  echo "$X rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2e_c2_n${N}_$X.json").read().strip().splitlines()[-1])
    print("$X N=$N ms_per_step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), d["losses_last_step"], d["config"]["parallelism"][:90])
except Exception as e:
    print("no line:", e)
PY
  grep -v "^\[W\|^$\|OMP_NUM\|^\*\*\*" gpurun_out/r2e_c2_n${N}_$X.log | tail -5
done
