"""Per-epilogue-mode time of k_spmm_light in the fused c2 training step (profile tags 0..3;
tag 0 = PLAIN x4 + FWD_FINAL, tag 3 = ADAM). Run with LGC_LIGHT_DEBUG=<bits> to switch parts off."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
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
    return tuple(torch.from_numpy(rng.integers(lo, hi, 1024)).to(dev)
                 for lo, hi in ((0, g.n_users), (g.n_users, g.num_nodes), (g.n_users, g.num_nodes)))
lib = _capi.lib()
for _ in range(3):
    tr.step(ei, ew, *batch(), 1e-4)
torch.cuda.synchronize()
n_tags, steps = 24, 10
ms, cnt = (C.c_double * n_tags)(), (C.c_longlong * n_tags)()
bs = [batch() for _ in range(steps)]
lib.lgc_profile_enable(1)
for b in bs:
    tr.step(ei, ew, *b, 1e-4)
torch.cuda.synchronize()
lib.lgc_profile_read(ms, cnt, n_tags)
lib.lgc_profile_enable(0)
print("LGC_LIGHT_DEBUG=%s" % os.environ.get("LGC_LIGHT_DEBUG", "0"),
      " light plain x4 + final: %.3f ms/step   light adam: %.3f   heavy: %.3f" %
      (ms[0] / steps, ms[3] / steps, sum(ms[4:8]) / steps))
