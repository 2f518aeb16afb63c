"""Golden checkpoint written by the REFERENCE'S OWN `save_model` (src/utils_v2.py:214-232).

Run in the build container only (needs /root/reference):

    python tests/golden/make_ref_checkpoint.py

Drives the reference `LightGCN` + `torch.optim.Adam` for the two mini-batches of tests/golden/tiny.npz
exactly like make_golden.py, saves the checkpoint with the reference's `save_model` (file layout
of `LightGCN_best.pt`), then runs ONE more reference step on a third mini-batch and stores the
weights / losses after it: a resume from the checkpoint must land there.
Writes tests/golden/ref_checkpoint_tiny.pt and ref_checkpoint_tiny_next.npz.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.reference_shim import load_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LR, DECAY = 0.005, 1e-4


def main():
    z = np.load(os.path.join(HERE, "tiny.npz"))
    ref_lightgcn, ref_utils = load_reference()
    n_users, n_items, dim, layers = int(z["n_users"]), int(z["n_items"]), int(z["dim"]), int(z["layers"])
    num_nodes = n_users + n_items
    frame = pd.DataFrame({"user_id_idx": z["user"], "item_id_idx": z["item"], "weight": z["weight"]})
    edge_index, edge_weight = ref_utils.df_to_graph(frame, True)
    model = ref_lightgcn.LightGCN(num_nodes, dim, layers)
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(z["init"]))
    optimizer = torch.optim.Adam(model.parameters(), LR)

    def step(u, p, n):
        optimizer.zero_grad()
        users, pos, neg = (torch.from_numpy(np.ascontiguousarray(a)) for a in (u, p, n))
        out = model(edge_index, ref_utils.batch_pos_neg_edges(users, pos, neg), edge_weight)
        size = len(users)
        bpr = model.recommendation_loss(out[:size], out[size:], 0) * size
        reg = ref_utils.regularization_loss(model.embedding.weight, size, users, pos, neg, DECAY)
        (bpr + reg).backward()
        optimizer.step()
        return [bpr.item(), reg.item(), (bpr + reg).item()]

    model.train()
    for t in z["triples"]:
        step(*t)
    assert np.array_equal(model.embedding.weight.detach().numpy(), z["f32_w2"])   # same run as tiny.npz
    path = os.path.join(HERE, "ref_checkpoint_tiny.pt")
    with contextlib.redirect_stdout(io.StringIO()):
        ref_utils.save_model(path, model, optimizer, 0.125, 0.25, epoch=2,
                             hyperparams={"latent_dim": dim, "n_layers": layers, "lr": LR, "decay": DECAY})
    third = z["triples"][0][:, ::-1].copy()                      # a third mini-batch: the first one reversed
    losses3 = step(*third)
    w3_f32 = model.embedding.weight.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "ref_checkpoint_tiny_next.npz"), triple=third,
                        losses=np.array(losses3, dtype=np.float64), w3=w3_f32)
    print(path, os.path.getsize(path), "bytes; losses of the resumed step", losses3)


if __name__ == "__main__":
    torch.manual_seed(0)
    main()
