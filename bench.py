#!/usr/bin/env python
"""Benchmark of the LightGCN hot path (BASELINE.json metric) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2]

A "step" is one training mini-batch of `TrainLightGCN.mini_batch_loop`
(reference `src/train_lightgcn.py:129-151`): K-layer forward over ALL edges, BPR + L2 loss,
K-layer backward, dense Adam -- on the synthetic Cosmetics-Shop-shaped graph c2
(1.6 M users x 54 K items, 5 M weighted edges -> nnz = 10 M directed, d = 64, K = 3, batch 1024).

metric `lgconv_gedges_per_s` = nnz * 2K * steps / seconds (directed edges through LGConv layer
passes, forward + backward). `value` has triples resident in HBM; `e2e` goes
through the public Python API with pinned host triples copied H2D and the three losses read
back D2H every step. Also reported: epoch seconds (122 steps), top-20 scoring users/s (c4),
the roofline of the dominant kernel, and the CPU port of the reference timed on this box.

`--impl reference` times the SAME step (K-layer forward, BPR + L2, backward, dense Adam on the same
graph and triples) with the reference's own CPU code on all host cores: the reference's `LightGCN`,
`BPRLoss` and `utils_v2` imported unmodified from the staged copy `oracle/_ref` (only the absent
third-party `LGConv` operator is the op-for-op restatement of `oracle/lgconv.py`; real PyG is not
installable here), else the oracle port. One CPU step at c2 is ~5 s on the box's 16 cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH, LR, DECAY = 1024, 0.005, 1e-4       # src/train_lightgcn.py:47-53


_REAL_STDOUT = None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- workload
def make_workload(config: str, n_batches: int):
    from gnn_ecommerce_b200 import synth
    t0 = time.time()
    n_users, n_items, n_edges, dim, layers = synth.CONFIGS[config]
    g = synth.make_graph(n_users, n_items, n_edges, seed=42)
    pl = synth.purchase_lists(g)
    rng = np.random.default_rng(44)
    triples = [synth.sample_triples(pl, BATCH, g.n_users, g.n_items, rng) for _ in range(n_batches)]
    bound = np.sqrt(6.0 / (g.num_nodes + dim))            # xavier_uniform_ (src/lightgcn.py:87)
    init = np.random.default_rng(43).uniform(-bound, bound, (g.num_nodes, dim)).astype(np.float32)
    log(f"[bench] synthetic {config}: N={g.num_nodes} nnz={2 * g.num_edges} d={dim} K={layers} "
        f"({time.time() - t0:.1f}s)")
    return g, dim, layers, init, triples


def algorithmic_bytes(g, dim_ld: int, layers: int, nnz: int):
    """int32 CSR, fp32 values. I = nnz*8 + (N+1)*4, T = N*ld*4,
    B_prop = K*I + (4K-1)*T, step = 2*B_prop + 7T."""
    n = g.num_nodes
    idx = nnz * 8 + (n + 1) * 4
    t = n * dim_ld * 4
    b_prop = layers * idx + (4 * layers - 1) * t
    return {"I": idx, "T": t, "B_prop": b_prop, "step": 2 * b_prop + 7 * t}


def workload_name(config: str, g, dim: int, layers: int) -> str:
    """Same string in both arms (the driver compares them)."""
    return (f"{config}: LightGCN K={layers} d={dim}, N={g.num_nodes}, nnz={2 * g.num_edges}, "
            f"batch={BATCH}, full training step (fwd+BPR+bwd+Adam)")


def stream_counts(layers: int):
    """Row streams of the epilogue per launch of a K-layer step (csrc/train_step.cu), SURVEY 8(d)
    accounting: forward K-1 x PLAIN (layer table written) + FWD_FINAL (K stored tables read, out
    written); backward K-1 x PLAIN + addend (read + write) + ADAM (addend, p, m, v read; p, m, v
    written)."""
    k = max(1, layers)
    return [1] * (k - 1) + [k + 1] + [2] * (k - 1) + [7]


def kernel_bytes(plan, num_nodes: int, ld: int, layers: int):
    """Algorithmic bytes of ONE launch of each SpMM kernel class, averaged over the 2K launches of a
    step: 8 B per edge (int32 source + fp32 weight), 4 B rowptr per row, every distinct gathered
    source row once (ld*4 B), plus the row streams of the fused epilogue (mode dependent)."""
    rb = ld * 4
    streams = stream_counts(layers)
    epi = rb * sum(streams) / float(len(streams))
    rows = plan.rows_edges * 8 + plan.rows_rows * 4 + plan.rows_sources * rb + plan.rows_rows * epi
    sweep = plan.sweep_edges * 8 + plan.sweep_rows * 4 + plan.sweep_sources * rb + plan.sweep_rows * epi
    return {"rows": rows, "sweep": sweep}


def light_kernel_bytes(g, graph, ld: int, layers: int = 3):
    """Round-1 kernels (no sweep plan for this row width): algorithmic bytes of one launch of
    k_spmm_light, averaged over the 2K launches of a K-layer step."""
    a = graph.arrays()
    rowptr = a["rowptr"].cpu().numpy().astype(np.int64)
    deg = np.diff(rowptr)
    light = deg <= graph.info.light_max_degree
    e_light = int(deg[light].sum())
    src = a["src"].cpu().numpy()
    rows_of_entry = np.repeat(np.arange(g.num_nodes), deg)
    distinct_src = int(np.unique(src[light[rows_of_entry]]).size)
    rb = ld * 4
    n_light = int(light.sum())
    gather = e_light * 8 + (g.num_nodes + 1) * 4 + distinct_src * rb
    streams = stream_counts(layers)
    epi = rb * sum(streams) / float(len(streams))
    return {"bytes_per_launch": gather + n_light * epi, "light_rows": n_light,
            "light_edges": e_light, "distinct_sources": distinct_src}


# ----------------------------------------------------------------------------- CPU legs
class CpuStepper:
    """One full reference training step per call on the host cores: the reference's own classes
    through `oracle/reference_shim.py` when its sources are available (kind "reference"), else the
    oracle port (kind "port")."""

    def __init__(self, g, dim, layers, init):
        import torch
        from oracle import port, reference_shim
        torch.set_num_threads(os.cpu_count() or 1)
        self.torch, self.port = torch, port
        self.ei, self.ew = port.df_to_graph(g.user, g.item, g.weight)
        self.ref_utils = None
        if reference_shim.reference_available():
            ref_lightgcn, self.ref_utils = reference_shim.load_reference()
            self.model = ref_lightgcn.LightGCN(g.num_nodes, dim, layers)
            self.kind = "reference"
            self.what = ("the reference's own LightGCN / BPRLoss / utils_v2 (staged unmodified under oracle/_ref) "
                         "around the op-for-op LGConv restatement (PyG itself is not installable here)")
            self._step = reference_shim.reference_train_step
        else:
            self.model = port.PortLightGCN(g.num_nodes, dim, layers)
            self.kind = "port"
            self.what = "oracle port of the reference's torch ops"
        with torch.no_grad():
            self.model.embedding.weight.copy_(torch.from_numpy(init))
        self.opt = torch.optim.Adam(self.model.parameters(), LR)
        self.cores = torch.get_num_threads()

    def step(self, triple):
        u, p, n = (self.torch.from_numpy(x) for x in triple)
        if self.ref_utils is not None:
            return self._step(self.ref_utils, self.model, self.opt, self.ei, self.ew, u, p, n, DECAY)
        return self.port.train_step(self.model, self.opt, self.ei, self.ew, u, p, n, DECAY)


def cpu_full_step(g, dim, layers, init, triple):
    """One full reference training step on the host cores -> (seconds, cores, kind, what)."""
    st = CpuStepper(g, dim, layers, init)
    t0 = time.perf_counter()
    st.step(triple)
    return time.perf_counter() - t0, st.cores, st.kind, st.what


def reference_arm(args):
    """`--impl reference`: the same full training step on the host CPU, all host threads, exactly
    `--warmup` + `--steps` steps (rank 0 only under torchrun)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_batches = args.steps + args.warmup
    g, dim, layers, init, triples = make_workload(args.config, n_batches)
    nnz = 2 * g.num_edges
    st = CpuStepper(g, dim, layers, init)
    for i in range(args.warmup):
        st.step(triples[i])
    t0 = time.perf_counter()
    for i in range(args.steps):
        losses = st.step(triples[args.warmup + i])
    dt = time.perf_counter() - t0
    value = nnz * 2 * layers * args.steps / dt / 1e9
    sample = (f"{args.steps} full training steps (K={layers} fwd + BPR + bwd + dense Adam) over the whole "
              f"{args.config} graph, {dt / args.steps:.2f} s per step; {st.what}")
    line = {"impl": "reference", "metric": "lgconv_gedges_per_s", "value": value, "unit": "GEdges/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak" if args.gpus == 1 else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.config, g, dim, layers)},
            "losses_last_step": [float(x) for x in losses],
            "cpu_baseline": {"value": value, "unit": "GEdges/s", "cores": st.cores, "kind": st.kind,
                             "sample": sample},
            "e2e": {"value": value, "unit": "GEdges/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c5"],
                    help="c5 (16 M users x 500 K items, 200 M interactions) is generated per rank: use --gpus 8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scoring", action="store_true")
    ap.add_argument("--no-epoch", action="store_true", help="skip the measured epoch (sampler + steps + eval)")
    ap.add_argument("--only-scoring", action="store_true", help="skip the training-step timing (dev aid)")
    ap.add_argument("--score-users", type=int, default=0, help="0 = all users (c4)")
    ap.add_argument("--c5-scale", type=float, default=1.0, help="shrink c5 by this factor (development aid)")
    ap.add_argument("--shard-mode", default="auto", choices=["auto", "bipartite", "rows"],
                    help="N > 1: how the training step is sharded")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1, bipartite sharding: item rows through the peer-memory kernel (lgc_item_exchange) "
                         "or through ncclAllReduce + lgc_epilogue_apply")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner
    # when NCCL_DEBUG is set) is sent to stderr; emit() writes to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from gnn_ecommerce_b200 import FusedBPRTrainer, LightGCN, _capi
    from gnn_ecommerce_b200.graph import padded_dim
    from gnn_ecommerce_b200.sharded import make_sharded_trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set: keep stdout to the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.lib()
    pk = peaks()
    if args.config == "c5":
        return bench_c5(args, world, rank, dev, lib, pk)

    n_batches = args.steps + args.warmup
    g, dim, layers, init, triples = make_workload(args.config, n_batches)
    nnz = 2 * g.num_edges
    ld = padded_dim(dim)

    ei = torch.from_numpy(g.edge_index()).to(dev)
    ew = torch.from_numpy(g.edge_weight()).to(dev)
    model, graph = None, None
    if world == 1:
        # N = 1: the drop-in module + fused single-GPU step
        model = LightGCN(g.num_nodes, dim, layers)
        with torch.no_grad():
            model.embedding.weight.copy_(torch.from_numpy(init))
        model = model.to(dev)
        trainer = FusedBPRTrainer(model, lr=LR)
        graph = model.graph(ei, ew)

        def step(u, p, n):
            return trainer.step(ei, ew, u, p, n, DECAY)

        def embedding():
            from gnn_ecommerce_b200 import ops
            with torch.no_grad():
                return ops.full_rows(model.get_embedding(ei, ew))
    else:
        # N > 1: the SAME graph is split over the ranks (strong scaling). Default "auto" = bipartite-
        # aware sharding (users partitioned, item table replicated, one all-reduce of the item
        # partial sums per layer); "rows" = destination rows partitioned + all-gather per layer
        # (SURVEY.md 8(e)). Both exchange the <= 3*batch loss rows with one small all-reduce.
        kw = {"exchange": args.exchange} if args.shard_mode != "rows" else {}
        trainer = make_sharded_trainer(ei, ew, g.num_nodes, dim, layers, torch.from_numpy(init),
                                       mode=args.shard_mode, lr=LR, **kw)
        shard_kind = type(trainer).__name__

        def step(u, p, n):
            return trainer.step(u, p, n, DECAY)

        def embedding():
            return trainer.embedding().contiguous()

    if args.only_scoring:
        sc = bench_scoring(embedding(), g, dev, args, pk, dim, world, rank)
        if rank == 0:
            emit({"scoring": sc})
        return finish(world, trainer)
    dev_triples = [tuple(torch.from_numpy(x).to(dev) for x in t) for t in triples]
    pin_triples = [tuple(torch.from_numpy(x).pin_memory() for x in t) for t in triples]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- device-resident timing: W warm-up + exactly K timed steps (inputs: 1 GB+ of tables,
    # far larger than the 126 MB L2, so no flush is needed between iterations)
    for i in range(args.warmup):
        step(*dev_triples[i])
    barrier()
    launches0 = lib.lgc_launch_count()
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        beg.record()
        for i in range(args.steps):
            loss3 = step(*dev_triples[args.warmup + i])
        end.record()
        barrier()
    launches = lib.lgc_launch_count() - launches0
    ms_total = max_over_ranks(beg.elapsed_time(end))
    ms_step = ms_total / args.steps
    last_losses = [float(x) for x in loss3.cpu().tolist()]
    gedges = nnz * 2 * layers * args.steps / (ms_total * 1e-3) / 1e9

    # ---- end to end through the public API: pinned host triples -> H2D, losses -> D2H per step
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        u, p, n = (x.to(dev, non_blocking=True) for x in pin_triples[args.warmup + i])
        host_losses = step(u, p, n).cpu()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_gedges = nnz * 2 * layers * args.steps / e2e_s / 1e9

    # ---- per-kernel-class durations (CUDA events on the launching stream) over K more steps
    n_tags = 24
    ms_arr, cnt_arr = (C.c_double * n_tags)(), (C.c_longlong * n_tags)()
    if world > 1:
        trainer.use_graph = False          # per-kernel events need eager launches
    eager0 = lib.lgc_launch_count()
    lib.lgc_profile_enable(1)
    for i in range(args.steps):
        step(*dev_triples[args.warmup + i])
    torch.cuda.synchronize()
    lib.lgc_profile_read(ms_arr, cnt_arr, n_tags)
    lib.lgc_profile_enable(0)
    if world > 1:
        if hasattr(trainer, "check_exchange"):
            trainer.check_exchange()       # a timed-out peer barrier invalidates the run: fail loudly
        trainer.use_graph = True
        launches = lib.lgc_launch_count() - eager0   # graph replays re-run these launches step for step
    light_ms = sum(ms_arr[t] for t in range(0, 4))
    light_cnt = sum(cnt_arr[t] for t in range(0, 4))
    heavy_ms = sum(ms_arr[t] for t in range(4, 8))
    finish_ms = sum(ms_arr[t] for t in range(8, 12))
    kernel_ms_total = sum(ms_arr)

    # ---- scoring (c4): all users x all items, top-20, user-sharded when N > 1
    score = None
    if not args.no_scoring:
        try:
            score = bench_scoring(embedding(), g, dev, args, pk, dim, world, rank)
        except Exception as e:  # a missing kernel must not hide the training number
            log(f"[bench] scoring failed: {e!r}")
            score = {"error": repr(e)}

    # ---- one measured epoch through the public API (single GPU): device sampler -> n_batch fused
    # steps -> top-20 of all evaluation users -> recall@20 (src/train_lightgcn.py:79-121,155-162)
    epoch = None
    if world == 1 and not args.no_epoch:
        try:
            epoch = bench_epoch(model, trainer, ei, ew, g, dev, dim)
        except Exception as e:
            log(f"[bench] measured epoch failed: {e!r}")
            epoch = {"error": repr(e)}

    if rank != 0:
        return finish(world, trainer)

    alg = algorithmic_bytes(g, ld, layers, nnz)
    light_avg_ms = light_ms / max(light_cnt, 1)
    class_ms = {"light": light_ms / args.steps, "heavy": heavy_ms / args.steps, "finish": finish_ms / args.steps,
                "bpr": ms_arr[12] / args.steps,
                "item_exchange": [round(ms_arr[14] / args.steps, 4), cnt_arr[14] // args.steps],   # incl. waiting for peers
                # per epilogue mode {plain, fwd-init, fwd-rmw (+ fwd-final), adam}: ms per step / launches per step
                "rows_by_mode": [[round(ms_arr[t] / args.steps, 4), cnt_arr[t] // args.steps] for t in range(0, 4)],
                "sweep_by_mode": [[round(ms_arr[t] / args.steps, 4), cnt_arr[t] // args.steps] for t in range(4, 8)]}
    step_gbs = alg["step"] / (ms_step * 1e-3) / 1e9
    if world == 1:
        plan = _capi.PlanInfo()
        _capi.check(lib.lgc_graph_plan_info(graph.handle, ld, C.byref(plan)), "lgc_graph_plan_info")
        traffic = None
        if plan.has_plan and plan.rows_rows > 0:
            # dominant kernel: k_spmm_rows (the low-degree rows: the users); the sweep is reported beside it
            kb = kernel_bytes(plan, g.num_nodes, ld, layers)
            kname, kbytes = "k_spmm_rows", kb["rows"]
            sweep_avg_ms = heavy_ms / max(sum(cnt_arr[t] for t in range(4, 8)), 1)
            extra = {"sweep_kernel": {"kernel": "k_spmm_sweep", "kernel_ms": sweep_avg_ms,
                                      "algorithmic_bytes_per_launch": kb["sweep"],
                                      "achieved": kb["sweep"] / (sweep_avg_ms * 1e-3) / 1e9,
                                      "frac": kb["sweep"] / (sweep_avg_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                                      "rows": int(plan.sweep_rows), "edges": int(plan.sweep_edges),
                                      "distinct_sources": int(plan.sweep_sources)},
                     "rows_kernel": {"rows": int(plan.rows_rows), "edges": int(plan.rows_edges),
                                     "distinct_sources": int(plan.rows_sources)}}
            prof_json = os.path.join(ROOT, "profiles", "ncu_rows_traffic.json")
        else:
            lk = light_kernel_bytes(g, graph, ld, layers)
            kname, kbytes, extra = "k_spmm_light", lk["bytes_per_launch"], {}
            prof_json = os.path.join(ROOT, "profiles", "ncu_light_traffic.json")
        if os.path.exists(prof_json) and args.config == "c2":
            try:   # DRAM bytes per launch of the same kernel from the committed ncu --set full capture (c2)
                traffic = float(json.load(open(prof_json))["dram_bytes_per_launch"])
            except Exception:
                traffic = None
        achieved = kbytes / (light_avg_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": pk["hbm_gbs"],
                    "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic,
                    "peak_source": pk["source"], "kernel_ms": light_avg_ms,
                    "kernel_share_of_step": light_ms / max(kernel_ms_total, 1e-9),
                    "algorithmic_bytes_per_launch": kbytes,
                    "step_algorithmic_bytes": alg["step"], "step_achieved_gbs": step_gbs,
                    "step_frac": step_gbs / pk["hbm_gbs"], "class_ms_per_step": class_ms, **extra}
    else:
        # whole step against N x the HBM peak (the collectives add NVLink time on top of it)
        if shard_kind == "BipartiteShardedTrainer":
            comm = 2 * layers * g.n_items * ld * 4 * 2 * (world - 1) // world     # ring all-reduce volume
            par = f"users partitioned over {world} GPUs, item table replicated, {exchange_kind(trainer)}"
        else:
            comm = (2 * layers - 1) * trainer.n_cols * ld * 4 * (world - 1) // world
            par = f"destination rows partitioned over {world} GPUs, NCCL all-gather per layer"
        roofline = {"bound": "hbm", "kernel": "whole step, all ranks", "achieved": step_gbs,
                    "peak": pk["hbm_gbs"] * world, "unit": "GB/s", "frac": step_gbs / (pk["hbm_gbs"] * world),
                    "traffic": None, "peak_source": pk["source"], "step_algorithmic_bytes": alg["step"],
                    "class_ms_per_step_rank0": class_ms, "nvlink_bytes_per_step_per_gpu": comm}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        dt, cores, kind, what = cpu_full_step(g, dim, layers, init, triples[0])
        cpu = {"value": nnz * 2 * layers / dt / 1e9, "unit": "GEdges/s", "cores": cores, "kind": kind,
               "sample": f"1 full training step (K={layers} fwd + BPR + bwd + dense Adam) at {args.config} "
                         f"shape, {dt:.1f} s; {what}",
               "step_s": dt}

    steps_per_epoch = int(g.num_edges / (BATCH * 40))        # src/train_lightgcn.py:92
    line = {
        "metric": "lgconv_gedges_per_s", "value": gedges, "unit": "GEdges/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, g, dim, layers),
                   "l2": "inputs (>=1 GB of tables per step) exceed the 126 MB L2; no flush",
                   "steps_per_epoch": steps_per_epoch,
                   "parallelism": "single GPU" if world == 1 else par},
        "epoch_s": ms_step * steps_per_epoch * 1e-3,
        "epoch_measured": epoch,
        "losses_last_step": last_losses,
        "e2e": {"value": e2e_gedges, "unit": "GEdges/s", "h2d_bytes_per_step": 3 * BATCH * 8,
                "d2h_bytes_per_step": 12, "ms_per_step": e2e_s / args.steps * 1e3,
                "losses_last_step": [float(x) for x in host_losses.tolist()]},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "scoring": score,
    }
    emit(line)
    return finish(world, trainer)


def bench_c5(args, world, rank, dev, lib, pk):
    """c5, the scale-up stress of BASELINE.json: 16 M users x 500 K items, 200 M weighted interactions,
    K = 3, d = 64, propagation + BPR step across the GPUs of one box. No process ever holds the global
    edge list: every rank generates ITS OWN users' interactions (same generator and weight law as c2,
    seeded per rank), builds its two rectangular operators from them, and the item degrees are the
    all-reduced partial degrees. Triples: every rank samples batch / world triples among its own
    purchasers (`batch_loader` semantics), all-gathered into the global batch.
    Parity at this size: 64 sampled user rows and 64 sampled item rows of the FIRST propagation layer
    against an fp64 numpy evaluation of the definition from the raw interactions (the item rows need
    every rank's edges: partial sums, all-reduced)."""
    import torch
    import torch.distributed as dist
    from gnn_ecommerce_b200 import synth
    from gnn_ecommerce_b200.graph import padded_dim
    from gnn_ecommerce_b200.sharded import BipartiteShardedTrainer, RowPartition
    n_users, n_items, n_edges, dim, layers = synth.CONFIGS["c5"]
    if args.c5_scale != 1.0:                      # a smaller stress of the same shape (development aid)
        n_users, n_items, n_edges = (int(x * args.c5_scale) for x in (n_users, n_items, n_edges))
    assert n_users % world == 0 and n_edges % world == 0
    nu, ne = n_users // world, n_edges // world
    t0 = time.time()
    g = synth.make_graph(nu, n_items, ne, seed=4200 + rank)          # this rank's users: local ids 0..nu-1
    log(f"[bench] rank {rank}: c5 shard {nu} users x {n_items} items, {ne} interactions ({time.time() - t0:.1f}s)")
    part = RowPartition(np.ones(n_users), world)
    part.bounds = np.arange(world + 1, dtype=np.int64) * nu
    part.max_rows = (nu + 3) // 4 * 4
    lo = rank * nu
    user = torch.from_numpy(g.user + lo).to(dev)
    item = torch.from_numpy(g.item - nu).to(dev)
    w = torch.from_numpy(g.weight).to(dev)
    bound = np.sqrt(6.0 / (n_users + n_items + dim))
    init_u = torch.from_numpy(np.random.default_rng(4300 + rank).uniform(-bound, bound, (nu, dim)).astype(np.float32))
    init_i = torch.from_numpy(np.random.default_rng(4299).uniform(-bound, bound, (n_items, dim)).astype(np.float32))
    trainer = BipartiteShardedTrainer.from_pairs(user, item, w, part, n_users, n_items, dim, layers, init_u, init_i,
                                                 lr=LR, exchange=args.exchange)
    ld = padded_dim(dim)
    nnz = 2 * n_edges

    # ---- parity of sampled rows of layer 1 (x1 = A_hat E0) against fp64 numpy from the raw interactions
    info = {}
    if True:
        b = trainer.backend
        x1_u = torch.zeros_like(trainer.e0_u)
        x1_i = torch.zeros_like(trainer.e0_i)
        b.spmm_ex(trainer.gu, ld, trainer.e0_i, trainer.ws_u, 0, y=x1_u, scale=1.0)
        b.spmm_ex(trainer.gi, ld, trainer.e0_u, trainer.ws_i, 0, y=x1_i, scale=1.0)
        if world > 1:
            dist.all_reduce(x1_i)
        gu_l, gi_l, gw = g.user.astype(np.int64), (g.item - nu).astype(np.int64), g.weight.astype(np.float64)
        # the fp32 deg^-1/2 the step itself uses (gcn_norm sums the weighted degree in fp32: for a hub item
        # with millions of edges that sum alone is ~1e-5 away from fp64, on one GPU as well): the check
        # isolates the propagation, w_hat = (dis[src] * w) * dis[dst] evaluated in fp64 from those
        dis_u = trainer.dis_u.double().cpu().numpy()
        dis_i = trainer.dis_i.double().cpu().numpy()
        what = dis_u[gu_l] * gw * dis_i[gi_l]
        rng = np.random.default_rng(99)
        su = rng.choice(nu, 64, replace=False)
        si = np.random.default_rng(98).choice(n_items, 64, replace=False)          # same items on every rank
        e0u, e0i = init_u.numpy().astype(np.float64), init_i.numpy().astype(np.float64)
        want_u = np.zeros((64, dim))
        for j, u_ in enumerate(su):
            m_ = gu_l == u_
            want_u[j] = (what[m_, None] * e0i[gi_l[m_]]).sum(0)
        part_i = np.zeros((64, dim))
        for j, i_ in enumerate(si):
            m_ = gi_l == i_
            part_i[j] = (what[m_, None] * e0u[gu_l[m_]]).sum(0)
        want_i = torch.from_numpy(part_i).to(dev)
        if world > 1:
            dist.all_reduce(want_i)
        got_u = x1_u[torch.from_numpy(su).to(dev), :dim].double().cpu().numpy()
        got_i = x1_i[torch.from_numpy(si).to(dev), :dim].double()
        err_u = float(np.abs(got_u - want_u).max() / np.abs(want_u).max())
        err_i = float((got_i - want_i).abs().max() / want_i.abs().max())
        info = {"layer1_sampled_rows": 128, "max_rel_err_user_rows": err_u, "max_rel_err_item_rows": err_i,
                "bar": 1e-5, "ok": bool(err_u < 1e-5 and err_i < 1e-5)}
        log(f"[bench] rank {rank}: c5 layer-1 parity of sampled rows: users {err_u:.2e}, items {err_i:.2e}")
        del x1_u, x1_i

    # ---- triples: batch / world per rank among its own purchasers, all-gathered
    pl = synth.purchase_lists(g)
    per = BATCH // world
    rng = np.random.default_rng(4400 + rank)
    n_batches = args.steps + args.warmup
    dev_triples = []
    for _ in range(n_batches):
        u_, p_, n_ = synth.sample_triples(pl, per, nu, n_items, rng)
        mine = torch.from_numpy(np.stack([u_ + lo, p_ - nu + n_users, n_ - nu + n_users])).to(dev)
        if world > 1:
            allt = torch.empty(world, 3, per, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allt, mine)
            mine = allt.permute(1, 0, 2).reshape(3, world * per)
        dev_triples.append((mine[0].contiguous(), mine[1].contiguous(), mine[2].contiguous()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        trainer.step(*dev_triples[i], DECAY)
    barrier()
    launches0 = lib.lgc_launch_count()
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index or 0) as clocks:
        beg.record()
        for i in range(args.steps):
            loss3 = trainer.step(*dev_triples[args.warmup + i], DECAY)
        end.record()
        barrier()
    ms_total = beg.elapsed_time(end)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    # end to end: host triples in, losses out, every step
    pin = [tuple(x.cpu().pin_memory() for x in t) for t in dev_triples]
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        u_, p_, n_ = (x.to(dev, non_blocking=True) for x in pin[args.warmup + i])
        host_losses = trainer.step(u_, p_, n_, DECAY).cpu()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    trainer.use_graph = False
    eager0 = lib.lgc_launch_count()
    n_tags = 24
    ms_arr, cnt_arr = (C.c_double * n_tags)(), (C.c_longlong * n_tags)()
    lib.lgc_profile_enable(1)                # one eager step with CUDA events around every launch of this library
    trainer.step(*dev_triples[0], DECAY)
    torch.cuda.synchronize()
    lib.lgc_profile_read(ms_arr, cnt_arr, n_tags)
    lib.lgc_profile_enable(0)
    launches = (lib.lgc_launch_count() - eager0) * args.steps
    class_ms = {"rows": [round(sum(ms_arr[0:4]), 4), int(sum(cnt_arr[0:4]))],
                "sweep": [round(sum(ms_arr[4:8]), 4), int(sum(cnt_arr[4:8]))],
                "finish": [round(sum(ms_arr[8:12]), 4), int(sum(cnt_arr[8:12]))],
                "bpr": [round(ms_arr[12], 4), int(cnt_arr[12])],
                "item_exchange": [round(ms_arr[14], 4), int(cnt_arr[14])]}      # [ms, launches] of one eager step
    trainer.use_graph = True
    trainer.check_exchange()
    if rank != 0:
        return finish(world, trainer)
    n_nodes = n_users + n_items
    idx = nnz * 8 + (n_nodes + 1) * 4
    tbytes = n_nodes * ld * 4
    step_bytes = 2 * (layers * idx + (4 * layers - 1) * tbytes) + 7 * tbytes
    gbs = step_bytes / (ms_step * 1e-3) / 1e9
    line = {
        "metric": "lgconv_gedges_per_s", "value": nnz * 2 * layers / (ms_step * 1e-3) / 1e9, "unit": "GEdges/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"c5: LightGCN K={layers} d={dim}, N={n_nodes}, nnz={nnz}, batch={BATCH}, "
                               f"full training step (fwd+BPR+bwd+Adam)",
                   "l2": "inputs (>= 4 GB of tables per rank and step) exceed the 126 MB L2; no flush",
                   "steps_per_epoch": int(n_edges / (BATCH * 40)),
                   "parallelism": f"users partitioned over {world} GPUs (each rank generated and holds only its own "
                                  f"users' interactions), item table replicated, {exchange_kind(trainer)}"},
        "epoch_s": ms_step * int(n_edges / (BATCH * 40)) * 1e-3,
        "losses_last_step": [float(x) for x in loss3.cpu().tolist()],
        "e2e": {"value": nnz * 2 * layers * args.steps / e2e_s / 1e9, "unit": "GEdges/s",
                "h2d_bytes_per_step": 3 * BATCH * 8, "d2h_bytes_per_step": 12, "ms_per_step": e2e_s / args.steps * 1e3,
                "losses_last_step": [float(x) for x in host_losses.tolist()]},
        "gpu_launches": int(launches), "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "kernel": "whole step, all ranks", "achieved": gbs, "peak": pk["hbm_gbs"] * world,
                     "unit": "GB/s", "frac": gbs / (pk["hbm_gbs"] * world), "traffic": None,
                     "peak_source": pk["source"], "step_algorithmic_bytes": step_bytes,
                     "class_ms_one_eager_step_rank0": class_ms,
                     "nvlink_bytes_per_step_per_gpu": 2 * layers * n_items * ld * 4 * 2 * (world - 1) // max(world, 1)},
        "parity": info, "cpu_baseline": None, "scoring": None}
    emit(line)
    return finish(world, trainer)


def exchange_kind(trainer) -> str:
    if getattr(trainer, "peer", None) is not None:
        return ("item partial sums reduced, epilogued and broadcast per layer by one peer-memory kernel over NVLink "
                "(lgc_item_exchange: P2P loads / stores, no NCCL on the item rows)")
    return "NCCL all-reduce of the item partial sums per layer + replicated epilogue (lgc_epilogue_apply)"


def finish(world, trainer):
    """Multi-rank teardown: drop the captured CUDA graph (it holds NCCL work) before the process
    group goes away, then leave without running interpreter-exit destructors in an arbitrary order."""
    if world > 1:
        import torch
        import torch.distributed as dist
        try:
            if hasattr(trainer, "close"):
                trainer.close()
            elif hasattr(trainer, "release_graph"):
                trainer.release_graph()
        except Exception as e:            # teardown must not turn a measured run into a failed one
            log(f"[bench] teardown: {e!r}")
        torch.cuda.synchronize()
        dist.barrier()
        sys.stderr.flush()
        os._exit(0)
    return 0


def bench_epoch(model, trainer, ei, ew, g, dev, dim):
    """ONE epoch as the reference runs it (`TrainLightGCN.train`, src/train_lightgcn.py:79-121): n_batch =
    int(E / (B * 40)) mini-batches drawn by the device sampler (`batch_loader`), each a fused step, then
    the evaluation of `test()` (:155-162): top-20 of every evaluation user and recall@20 -- one timed
    region (CUDA events around it; the only host reads are the epoch's mean losses and the two metrics)."""
    import torch
    from gnn_ecommerce_b200 import ops, scoring, synth
    from gnn_ecommerce_b200.sampler import DeviceSampler
    k = 20
    held = synth.make_heldout(g, max(1000, g.n_users // 50))        # users with a held-out purchase (~2 %)
    pl = synth.purchase_lists(g, held)
    sampler = DeviceSampler.from_lists(pl.users, pl.pos_ptr, pl.pos_items, pl.ign_ptr, pl.ign_items,
                                       g.n_users, g.n_items, dev, seed=7)
    ptr, items = synth.seen_lists(g, held.users)
    seen = scoring.SeenLists.from_numpy(ptr, items, dev)
    eval_users = torch.from_numpy(held.users).to(dev)
    held_ptr, held_items = torch.from_numpy(held.ptr).to(dev), torch.from_numpy(held.items).to(dev)
    n_batch = int(g.num_edges / (BATCH * 40))

    def run():
        acc = torch.zeros(3, dtype=torch.float32, device=dev)
        for _ in range(n_batch):
            u, p, n = sampler.sample(BATCH)
            acc += trainer.step(ei, ew, u, p, n, DECAY)
        with torch.no_grad():
            rows = ops.full_rows(model.cached_embedding(ei, ew))
            top, _ = scoring.score_topk(rows[:g.n_users], rows[g.n_users:], eval_users, seen.ptr, seen.items, k, d=dim)
            prec, rec, _ = scoring.mark_mapk(top, held_ptr, held_items)
        return (acc / n_batch).cpu().tolist(), prec, rec

    run()                                                           # warm-up epoch (allocations, schedules)
    torch.cuda.synchronize()
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    beg.record()
    losses, prec, rec = run()
    end.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    return {"epoch_s": beg.elapsed_time(end) * 1e-3, "wall_s": wall, "n_batch": n_batch, "batch": BATCH,
            "eval_users": int(held.users.size), "k": k, "mean_losses": losses, "precision_at_k": prec,
            "recall_at_k": rec,
            "what": "device sampler + n_batch fused steps + top-20 of the evaluation users + recall@20, one timed region"}


def bench_scoring(rows, g, dev, args, pk, dim, world=1, rank=0):
    """c4: top-20 of all users x all items from the final embedding table `rows` ([N, >=dim]);
    users are sharded over the ranks (no communication), time = max over ranks."""
    import torch
    import torch.distributed as dist
    from gnn_ecommerce_b200 import _capi, scoring, synth
    from gnn_ecommerce_b200.sharded import shard_users
    k = 20
    n_score = args.score_users or g.n_users
    u0, u1 = shard_users(n_score, world, rank)
    users = torch.arange(u0, u1, device=dev)
    ptr, items = synth.seen_lists(g, np.arange(u0, u1))
    seen = scoring.SeenLists.from_numpy(ptr, items, dev)
    user_t, item_t = rows[:g.n_users], rows[g.n_users:]

    out = None

    def run():
        # the result tensors of the first call are reused: no allocation inside the timed calls
        return scoring.score_topk(user_t, item_t, users, seen.ptr, seen.items, k, d=dim, return_stats=True, out=out)
    for _ in range(2):
        top, sc, stats = run()
        out = (top, sc)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    rep_ms = []
    for _ in range(reps):
        beg.record()
        top, sc, stats = run()
        end.record()
        torch.cuda.synchronize()
        rep_ms.append(beg.elapsed_time(end))
    ms = float(np.median(rep_ms))
    n_tags = 24
    ms_arr, cnt_arr = (C.c_double * n_tags)(), (C.c_longlong * n_tags)()
    lib = _capi.lib()
    lib.lgc_profile_enable(1)
    run()
    torch.cuda.synchronize()
    lib.lgc_profile_read(ms_arr, cnt_arr, n_tags)
    lib.lgc_profile_enable(0)
    names = {16: "convert", 17: "gemm", 18: "threshold", 22: "scan", 19: "rescore", 20: "select",
             21: "exhaustive"}
    class_ms = {names[t]: ms_arr[t] for t in names}
    gemm_ms = max(ms_arr[17], 1e-9)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # ---- end to end through the public call: pinned host user ids in, host top-k array out
    host_users = torch.arange(u0, u1, dtype=torch.int64).pin_memory()
    e2e_ms = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        uid = host_users.to(dev, non_blocking=True)
        top_e, _ = scoring.score_topk(user_t, item_t, uid, seen.ptr, seen.items, k, d=dim, out=out)
        host_top = scoring.topk_to_host(top_e)
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    e2e = float(np.median(e2e_ms))
    if world > 1:
        t = torch.tensor([e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = float(t.item())
    # ---- the timed result itself is checked: 512 sampled users against a plain fp32 torch product
    # with the reference's multiplicative mask (src/lightgcn.py:173-177); ties within 2e-6 may swap
    n_chk = min(512, u1 - u0)
    chk = torch.from_numpy(np.random.default_rng(5).choice(u1 - u0, n_chk, replace=False)).to(dev)
    with torch.no_grad():
        old_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        pred = user_t[users[chk], :dim].double() @ item_t[:, :dim].double().t()
        torch.backends.cuda.matmul.allow_tf32 = old_tf32
        mask = torch.zeros_like(pred)
        lo, hi = seen.ptr[chk], seen.ptr[chk + 1]
        for i in range(n_chk):
            mask[i, seen.items[int(lo[i]):int(hi[i])]] = 1.0
        ref = pred * (1 - mask)
        ref_kth = ref.topk(k, dim=-1).values[:, -1]
        got = top[chk]
        invalid = int(((got < 0) | (got >= g.n_items)).any(dim=-1).sum())      # ids outside [0, n_items)
        got_scores = torch.gather(ref, 1, got.clamp(0, g.n_items - 1))
        tol = 2e-6 * ref.abs().amax(dim=-1)
        bad = int(((got_scores.min(dim=-1).values + tol) < ref_kth).sum())
        dup = int((torch.sort(top[chk], dim=-1).values.diff(dim=-1) == 0).any(dim=-1).sum())
    check = {"users_checked": n_chk, "users_with_a_wrong_item": bad, "users_with_duplicates": dup,
             "users_with_invalid_ids": invalid,
             "host_array_matches_device": bool((torch.from_numpy(host_top.astype(np.int64)).to(dev) == top_e).all())}
    flops_local = 2.0 * (u1 - u0) * g.n_items * dim
    flops = 2.0 * n_score * g.n_items * dim
    tf = flops / (ms * 1e-3) / 1e12
    st = stats.cpu().tolist()
    return {"metric": "top20_users_per_s", "value": n_score / (ms * 1e-3), "unit": "users/s",
            "users": n_score, "items": g.n_items, "k": k, "ms": ms, "n_gpus": world,
            "roofline": {"bound": "tensor", "kernel": "k_score_gemm",
                         "achieved": flops_local / (gemm_ms * 1e-3) / 1e12,
                         "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": flops_local / (gemm_ms * 1e-3) / 1e12 / pk["bf16_tflops"], "traffic": None,
                         "peak_source": pk["source"], "kernel_ms": gemm_ms,
                         "whole_call_tflops": tf, "whole_call_frac": tf / (pk["bf16_tflops"] * world)},
            "class_ms": class_ms, "rep_ms": rep_ms,
            "e2e": {"value": n_score / (e2e * 1e-3), "unit": "users/s", "ms": e2e, "rep_ms": e2e_ms,
                    "h2d_bytes": int((u1 - u0) * 8), "d2h_bytes": int((u1 - u0) * k * 4),
                    "what": "pinned host user ids -> device, top-k, int32 ids -> pinned host array"},
            "check": check,
            "fallback_users": st[0], "candidate_groups": st[1]}


if __name__ == "__main__":
    sys.exit(main())
