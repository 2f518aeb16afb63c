"""Per-phase cycles of k_spmm_light over a few fused training steps at c2 (LGC_LIGHT_PHASES=1).
Run: LGC_LIGHT_PHASES=1 [LGC_LIGHT_BALANCED=0|1] python profiles/experiments/light_phases.py"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("LGC_LIGHT_PHASES", "1")
import numpy as np
import torch
from gnn_ecommerce_b200 import FusedBPRTrainer, LightGCN, _capi, synth
from oracle import port

dev = torch.device("cuda:0")
g = synth.make_graph(1_600_000, 54_000, 5_000_000, seed=42)
ei, ew = port.df_to_graph(g.user, g.item, g.weight)
ei, ew = ei.to(dev), ew.to(dev)
model = LightGCN(g.num_nodes, 64, 3).to(dev)
tr = FusedBPRTrainer(model, lr=0.005)
rng = np.random.default_rng(0)
def batch():
    u = torch.from_numpy(rng.integers(0, g.n_users, 1024)).to(dev)
    p = torch.from_numpy(rng.integers(g.n_users, g.num_nodes, 1024)).to(dev)
    n = torch.from_numpy(rng.integers(g.n_users, g.num_nodes, 1024)).to(dev)
    return u, p, n
lib = _capi.lib()
out = (C.c_ulonglong * 8)()
for _ in range(3):
    tr.step(ei, ew, *batch(), 1e-4)
lib.lgc_debug_light_phases(out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
bs = [batch() for _ in range(5)]
e0.record()
for b in bs:
    tr.step(ei, ew, *b, 1e-4)
e1.record(); torch.cuda.synchronize()
lib.lgc_debug_light_phases(out)
v = list(out)
tiles = max(v[5], 1)
names = ["wait csr", "gathers", "wait operands", "epilogue", "store+refill"]
tot = sum(v[:5])
print(f"balanced={os.environ.get('LGC_LIGHT_BALANCED', '1')}  ms/step {e0.elapsed_time(e1) / 5:.3f}  warp-tiles {tiles}")
for n, c in zip(names, v[:5]):
    print(f"  {n:14s} {c / tiles:8.0f} cycles/tile  {100 * c / tot:5.1f} %")
print(f"  total          {tot / tiles:8.0f}")
