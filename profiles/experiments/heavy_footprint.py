"""Experiment (round 1f): how does k_spmm_heavy's time depend on the footprint of the gathered
table?  Item rows of the c2 graph gather user rows; the user ids are folded into C columns
(u % C), so the edge count and the degree distribution stay fixed while the table shrinks from
410 MB (C = 1.6 M) to 13 MB.  If the time drops sharply once the table fits in L2, blocking the
hub rows by source range (column blocking) pays.  Run: python profiles/experiments/heavy_footprint.py"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from gnn_ecommerce_b200 import _capi, synth
from gnn_ecommerce_b200.sharded import CudaBackend

dev = torch.device("cuda:0")
g = synth.make_graph(1_600_000, 54_000, 5_000_000, seed=42)
be, lib = CudaBackend(), _capi.lib()
item = torch.from_numpy(g.item - g.n_users).to(dev)
user = torch.from_numpy(g.user).to(dev)
w = torch.rand(item.numel(), device=dev)
n_tags = 24
for cols in (1_600_000, 800_000, 400_000, 200_000, 100_000, 50_000):
    for shuffle in (False, True):
        src = user % cols
        if shuffle:                                    # same footprint, sources decorrelated from the id order
            perm = torch.randperm(cols, device=dev)
            src = perm[src]
        h = be.build_rect(src, item, w, g.n_items, cols)
        x = torch.randn(cols, 64, device=dev)
        y = torch.empty(g.n_items, 64, device=dev)
        ws = be.workspace(h, 64, dev)
        for _ in range(3):
            be.spmm_ex(h, 64, x, ws, 0, y=y)
        torch.cuda.synchronize()
        ms_arr, cnt_arr = (C.c_double * n_tags)(), (C.c_longlong * n_tags)()
        lib.lgc_profile_enable(1)
        for _ in range(10):
            be.spmm_ex(h, 64, x, ws, 0, y=y)
        torch.cuda.synchronize()
        lib.lgc_profile_read(ms_arr, cnt_arr, n_tags)
        lib.lgc_profile_enable(0)
        light = sum(ms_arr[0:4]) / 10; heavy = sum(ms_arr[4:8]) / 10; fin = sum(ms_arr[8:12]) / 10
        print(f"cols {cols:8d} ({cols*256/1e6:6.1f} MB) shuffle {int(shuffle)}: light {light:.4f} heavy {heavy:.4f} finish {fin:.4f} ms", flush=True)
        be.destroy(h)
