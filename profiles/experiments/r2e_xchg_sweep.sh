#!/bin/bash
# r2e_xchg_sweep.sh N: c2 step time for several CTA caps of lgc_item_exchange, and the NCCL path -> stdout
N=${1:-2}
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 --config c2 --no-scoring --exchange $1 > gpurun_out/r2e_sweep.json 2> gpurun_out/r2e_sweep.log
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2e_sweep.json").read().strip().splitlines()[-1])
    print("$1 ctas=${LGC_XCHG_CTAS:-default} N=$N ms_per_step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("no line:", e)
PY
}
for c in 32 64 148 296; do LGC_XCHG_CTAS=$c run peer; done
run nccl
