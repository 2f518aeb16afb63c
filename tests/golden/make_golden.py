"""Generate the committed golden vectors from the REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports `/root/reference/src/lightgcn.py` and `utils_v2.py` unmodified through
`oracle/reference_shim.py` (the one absent third-party operator, PyG `LGConv`, is the op-for-op
restatement in `oracle/lgconv.py`), drives them exactly like `TrainLightGCN.mini_batch_loop`
(`src/train_lightgcn.py:123-153`) and `TrainLightGCN.test` (`:155-162`) on seeded synthetic
inputs, and writes `tests/golden/{tiny,c1}.npz`. The reference holds no tests or golden vectors
of its own, so these files are what pins the oracle and the CUDA path.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gnn_ecommerce_b200 import synth  # noqa: E402
from oracle.reference_shim import load_reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LR, DECAY, BATCH = 0.005, 1e-4, 1024     # src/train_lightgcn.py:47-53


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def init_weight(num_nodes: int, dim: int, seed: int) -> np.ndarray:
    """Same law as `xavier_uniform_` (src/lightgcn.py:87) but from numpy's PCG64 so the
    fixture does not depend on torch's RNG stream."""
    bound = np.sqrt(6.0 / (num_nodes + dim))
    return np.random.default_rng(seed).uniform(-bound, bound, (num_nodes, dim)).astype(np.float32)


def frame_of(g: synth.BipartiteGraph) -> pd.DataFrame:
    return pd.DataFrame({"user_id_idx": g.user, "item_id_idx": g.item, "weight": g.weight})


def run_reference(g, dim, layers, triples, eval_users, seen_ptr, seen_items, heldout_lists,
                  k, dtype, num_nodes=None):
    ref_lightgcn, ref_utils = load_reference()
    num_nodes = num_nodes or g.num_nodes
    edge_index, edge_weight = ref_utils.df_to_graph(frame_of(g), True)
    model = ref_lightgcn.LightGCN(num_nodes, dim, layers)
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(init_weight(num_nodes, dim, 43)))
    if dtype == torch.float64:
        model = model.double()
        edge_weight = edge_weight.double()
    optimizer = torch.optim.Adam(model.parameters(), LR)      # src/train_lightgcn.py:58
    res = {}
    with torch.no_grad():
        res["out0"] = model.get_embedding(edge_index, edge_weight).numpy().copy()
    losses, grads, weights, scores = [], [], [], []
    model.train()
    for (u, p, n) in triples:                                  # src/train_lightgcn.py:129-151
        optimizer.zero_grad()
        users, pos, neg = torch.from_numpy(u), torch.from_numpy(p), torch.from_numpy(n)
        labels = ref_utils.batch_pos_neg_edges(users, pos, neg)
        out = model(edge_index, labels, edge_weight)
        size = len(users)
        bpr = model.recommendation_loss(out[:size], out[size:], 0) * size
        reg = ref_utils.regularization_loss(model.embedding.weight, size, users, pos, neg, DECAY)
        loss = bpr + reg
        loss.backward()
        grads.append(model.embedding.weight.grad.numpy().copy())
        optimizer.step()
        scores.append(out.detach().numpy().copy())
        losses.append([bpr.item(), reg.item(), loss.item()])
        weights.append(model.embedding.weight.detach().numpy().copy())
    res.update(losses=np.array(losses, dtype=np.float64), grads=grads, weights=weights,
               scores=scores)
    model.eval()
    if eval_users is not None:
        from oracle.port import dense_seen_mask
        mask = dense_seen_mask(seen_ptr, seen_items, g.n_items)
        with torch.no_grad():
            top_df = model.recommendK(edge_index, edge_weight, g.n_users, g.n_items, mask,
                                      list(eval_users), k)          # src/train_lightgcn.py:159
            pos_df = pd.DataFrame({"user_id_idx": list(eval_users),
                                   "item_id_idx_list": heldout_lists})
            prec, rec, _ = model.MARK_MAPK(pos_df, top_df, k)
            emb = model.get_embedding(edge_index, edge_weight)
            src, dst = torch.split(emb, [g.n_users, g.n_items])
            masked = (src[list(eval_users)] @ dst.t()) * (1 - mask.to(emb.dtype))
        res.update(topk=np.array(top_df["top_rlvnt_itm"].tolist(), dtype=np.int64),
                   precision=float(prec), recall=float(rec),
                   topk_scores=np.take_along_axis(
                       masked.numpy(), np.array(top_df["top_rlvnt_itm"].tolist()), axis=1),
                   kth_gap=(masked.topk(k + 1, dim=-1).values[:, k - 1]
                            - masked.topk(k + 1, dim=-1).values[:, k]).numpy())
    res["edge_index"], res["edge_weight"] = edge_index.numpy(), edge_weight.numpy()
    return res


def eval_inputs(g, n_eval, seed):
    held = synth.make_heldout(g, n_eval, seed=seed)
    ptr, items = synth.seen_lists(g, held.users)
    lists = [held.items[held.ptr[i]:held.ptr[i + 1]].tolist() for i in range(held.users.size)]
    return held, ptr, items, lists


def make_tiny():
    """Small enough to store whole; has isolated nodes (num_nodes > last id + 1), duplicate
    triples and a user whose every positive score is negative (multiplicative-mask quirk)."""
    g = synth.make_graph(60, 25, 300, seed=7)
    num_nodes = g.num_nodes + 3                    # three isolated nodes at the end
    dim, layers, k = 16, 2, 5
    rng = np.random.default_rng(44)
    pl = synth.purchase_lists(g)
    triples = []
    for _ in range(2):
        u, p, n = synth.sample_triples(pl, min(16, pl.users.size), g.n_users, g.n_items, rng)
        u, p, n = (np.concatenate([a, a[:4]]) for a in (u, p, n))     # duplicates
        triples.append((u, p, n))
    held, ptr, items, lists = eval_inputs(g, 20, 45)
    out = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        # recommendK splits [n_users, n_items]: evaluate on the graph without isolated tail
        r = run_reference(g, dim, layers, triples, held.users, ptr, items, lists, k, dt)
        riso = run_reference(g, dim, layers, triples, None, None, None, None, k, dt, num_nodes)
        out.update({f"{tag}_out0": r["out0"], f"{tag}_losses": r["losses"],
                    f"{tag}_grad0": r["grads"][0], f"{tag}_w1": r["weights"][0],
                    f"{tag}_w2": r["weights"][1], f"{tag}_scores0": r["scores"][0],
                    f"{tag}_topk": r["topk"], f"{tag}_topk_scores": r["topk_scores"],
                    f"{tag}_precision": r["precision"], f"{tag}_recall": r["recall"],
                    f"{tag}_iso_out0": riso["out0"], f"{tag}_iso_w2": riso["weights"][1],
                    f"{tag}_iso_losses": riso["losses"]})
    np.savez_compressed(
        os.path.join(HERE, "tiny.npz"), n_users=g.n_users, n_items=g.n_items,
        num_nodes_iso=num_nodes, dim=dim, layers=layers, k=k, user=g.user, item=g.item,
        weight=g.weight, edge_index=r["edge_index"], edge_weight=r["edge_weight"],
        init=init_weight(g.num_nodes, dim, 43), init_iso=init_weight(num_nodes, dim, 43),
        triples=np.array([np.stack(t) for t in triples]), eval_users=held.users,
        seen_ptr=ptr, seen_items=items, held_ptr=held.ptr, held_items=held.items, **out)


def make_c1():
    g = synth.make_config_graph("c1", seed=42)
    _, _, _, dim, layers = synth.CONFIGS["c1"]
    k = 20
    held, ptr, items, lists = eval_inputs(g, 256, 45)
    pl = synth.purchase_lists(g, held)
    rng = np.random.default_rng(44)
    triples = [synth.sample_triples(pl, BATCH, g.n_users, g.n_items, rng) for _ in range(2)]
    rows = np.sort(np.random.default_rng(46).choice(g.num_nodes, 384, replace=False))
    touched = np.unique(np.concatenate(triples[0]))[:128]
    rows = np.unique(np.concatenate([rows, touched]))
    out = {}
    err = {}
    res = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        r = res[tag] = run_reference(g, dim, layers, triples, held.users, ptr, items, lists, k, dt)
        out.update({f"{tag}_out0_rows": r["out0"][rows], f"{tag}_losses": r["losses"],
                    f"{tag}_grad0_rows": r["grads"][0][rows],
                    f"{tag}_w1_rows": r["weights"][0][rows],
                    f"{tag}_w2_rows": r["weights"][1][rows], f"{tag}_scores0": r["scores"][0],
                    f"{tag}_topk": r["topk"], f"{tag}_topk_scores": r["topk_scores"],
                    f"{tag}_kth_gap": r["kth_gap"],
                    f"{tag}_precision": r["precision"], f"{tag}_recall": r["recall"]})
    # error budget of the reference's own fp32 run against its fp64 run (max-norm relative)
    for name, a, b in (("out0", res["f32"]["out0"], res["f64"]["out0"]),
                       ("grad0", res["f32"]["grads"][0], res["f64"]["grads"][0]),
                       ("w1", res["f32"]["weights"][0], res["f64"]["weights"][0]),
                       ("w2", res["f32"]["weights"][1], res["f64"]["weights"][1])):
        err[f"budget_{name}"] = float(np.abs(a - b).max() / np.abs(b).max())
    from oracle.port import csr_by_target
    csr = csr_by_target(res["f32"]["edge_index"], res["f32"]["edge_weight"], g.num_nodes)
    np.savez_compressed(
        os.path.join(HERE, "c1.npz"), n_users=g.n_users, n_items=g.n_items, dim=dim,
        layers=layers, k=k, graph_seed=42, init_seed=43,
        sha_edge_index=sha(res["f32"]["edge_index"]), sha_edge_weight=sha(res["f32"]["edge_weight"]),
        sha_init=sha(init_weight(g.num_nodes, dim, 43)),
        sha_rowptr=sha(csr["rowptr"]), sha_src=sha(csr["src"]), sha_w_hat_csr=sha(csr["w_hat_csr"]),
        deg=csr["deg"], dis=csr["dis"], count_deg=csr["count_deg"].astype(np.int32),
        triples=np.array([np.stack(t) for t in triples]), rows=rows, eval_users=held.users,
        seen_ptr=ptr, seen_items=items, held_ptr=held.ptr, held_items=held.items, **out, **err)
    print("c1 budgets:", err, "losses:", res["f32"]["losses"].tolist(),
          "recall", res["f32"]["recall"])


if __name__ == "__main__":
    torch.manual_seed(0)
    make_tiny()
    make_c1()
    for f in ("tiny.npz", "c1.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
