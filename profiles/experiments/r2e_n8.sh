#!/bin/bash
# r2e_n8.sh: the driver's own N=8 command (c2, scoring included) with the peer-memory exchange, then the NCCL path
N=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2e_c2_n8_peer.json 2> gpurun_out/r2e_c2_n8_peer.log; echo "peer rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --no-scoring --exchange nccl > gpurun_out/r2e_c2_n8_nccl.json 2> gpurun_out/r2e_c2_n8_nccl.log; echo "nccl rc=$?"
python - <<PY
import json
for x in ("peer", "nccl"):
    try:
        d = json.loads(open(f"gpurun_out/r2e_c2_n8_{x}.json").read().strip().splitlines()[-1])
        sc = d.get("scoring") or {}
        print(x, "ms_per_step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), d["losses_last_step"],
              d["roofline"].get("class_ms_per_step_rank0"), "scoring ms", sc.get("ms"), sc.get("check"))
    except Exception as e:
        print(x, "no line:", e)
PY
grep -v "^\[W\|^$\|OMP_NUM\|^\*\*\*\|Assertion" gpurun_out/r2e_c2_n8_peer.log | tail -6
