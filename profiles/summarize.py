#!/usr/bin/env python
"""Turn an Nsight Compute report into the compact JSON summaries committed under profiles/.

    python profiles/summarize.py gpurun_out/r1d_step.ncu-rep profiles/r1d_ncu_full_step.json "<command that was profiled>"

Reads `ncu -i <rep> --page raw --csv` (works without a GPU) and keeps, per captured launch, the
metrics DESIGN.md / bench.py quote: duration, DRAM bytes, achieved occupancy, registers, L1/L2 hit
rates, issue utilisation, tensor-pipe activity and the top warp-stall reasons."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[col["Kernel Name"]].replace("void ", "")[:90]}
        for k in KEEP:
            if k in col and r[col[k]] != "":
                d[k] = f"{r[col[k]]} {units[col[k]]}".strip()
        stalls = []
        for h, i in col.items():
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "0"):
                stalls.append((float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]))
        d["top_stalls_per_issue"] = {n: round(v, 3) for v, n in sorted(stalls, reverse=True)[:4]}
        launches.append(d)
    json.dump({"command": cmd, "report": rep.split("/")[-1], "launches": launches}, open(out, "w"), indent=1)
    print(f"{out}: {len(launches)} launches")


if __name__ == "__main__":
    main()
