#!/bin/bash
# r2_scale.sh N CONFIG [extra bench args]: one multi-rank bench line -> gpurun_out/r2_scale_<config>_n<N>.json
N=$1; CFG=$2; shift; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --config $CFG "$@" > gpurun_out/r2_scale_${CFG}_n$N.json 2> gpurun_out/r2_scale_${CFG}_n$N.log
echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_scale_${CFG}_n$N.json").read().strip().splitlines()[-1])
sc = d.get("scoring") or {}
print("$CFG N=$N ms_per_step", round(d["ms_per_step"], 4), "GEdges/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2),
      d["roofline"].get("class_ms_per_step_rank0"), "scoring ms", sc.get("ms"), "users/s", sc.get("value"))
PY
