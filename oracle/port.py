"""CPU restatement of the reference's LightGCN hot path (TEST INFRASTRUCTURE, kind = "port").

This is the oracle that travels to the GPU box (where `/root/reference` does not exist). Each
function cites the reference lines it follows; `tests/test_oracle.py` checks it against the
golden vectors the reference's own code produced (`tests/golden/make_golden.py`) and, in the
build container, bit-for-bit against the reference imported through `oracle/reference_shim.py`.

Works in fp32 (the reference's dtype) or fp64 (for error budgeting: `dtype=torch.float64`).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

from oracle.lgconv import LGConv


# --------------------------------------------------------------------------- graph / labels
def df_to_graph(user: np.ndarray, item: np.ndarray, weight: np.ndarray) -> Tuple[Tensor, Tensor]:
    """`df_to_graph(train_df, True)` (reference `src/utils_v2.py:146-165`) on plain arrays:
    `item` is already offset by n_users (`src/utils_v2.py:128`)."""
    u_t = torch.from_numpy(np.ascontiguousarray(user, dtype=np.int64))
    i_t = torch.from_numpy(np.ascontiguousarray(item, dtype=np.int64))
    w_t = torch.from_numpy(np.ascontiguousarray(weight, dtype=np.float32))
    edge_index = torch.stack((torch.cat([u_t, i_t]), torch.cat([i_t, u_t])))
    return edge_index, torch.cat([w_t, w_t])


def batch_pos_neg_edges(users: Tensor, pos: Tensor, neg: Tensor) -> Tensor:
    """[[u;u],[p;n]] label pairs (reference `src/utils_v2.py:184-190`)."""
    return torch.stack((torch.cat([users, users]), torch.cat([pos, neg])))


# --------------------------------------------------------------------------- model
class PortLightGCN(torch.nn.Module):
    """Same state (`embedding.weight` [N,d], buffer `alpha` [K+1]) and the same arithmetic as
    the reference's `LightGCN` (`src/lightgcn.py:58-125`), written independently."""

    def __init__(self, num_nodes: int, embedding_dim: int, num_layers: int,
                 alpha=None, dtype: torch.dtype = torch.float32):
        super().__init__()
        self.num_nodes, self.embedding_dim, self.num_layers = num_nodes, embedding_dim, num_layers
        if alpha is None:
            alpha = 1.0 / (num_layers + 1)                         # src/lightgcn.py:72-73
        if not isinstance(alpha, Tensor):
            alpha = torch.tensor([alpha] * (num_layers + 1))       # src/lightgcn.py:78
        assert alpha.numel() == num_layers + 1
        self.register_buffer("alpha", alpha.to(dtype))
        self.embedding = torch.nn.Embedding(num_nodes, embedding_dim, dtype=dtype)
        torch.nn.init.xavier_uniform_(self.embedding.weight)       # src/lightgcn.py:87
        self.conv = LGConv()

    def get_embedding(self, edge_index: Tensor, edge_weight: Optional[Tensor]) -> Tensor:
        """out = sum_l alpha_l * A_hat^l E0 as a running sum (src/lightgcn.py:91-99)."""
        if edge_weight is not None:
            edge_weight = edge_weight.to(self.embedding.weight.dtype)
        x = self.embedding.weight
        out = x * self.alpha[0]
        for layer in range(self.num_layers):
            x = self.conv(x, edge_index, edge_weight)
            out = out + x * self.alpha[layer + 1]
        return out

    def forward(self, edge_index: Tensor, edge_label_index: Optional[Tensor] = None,
                edge_weight: Optional[Tensor] = None) -> Tensor:
        """Row-wise dot products of the final embeddings (src/lightgcn.py:101-125)."""
        if edge_label_index is None:
            edge_label_index = edge_index
        out = self.get_embedding(edge_index, edge_weight)
        return (out[edge_label_index[0]] * out[edge_label_index[1]]).sum(dim=-1)


def bpr_loss(pos_rank: Tensor, neg_rank: Tensor) -> Tensor:
    """`recommendation_loss(pos, neg, 0) * size` exactly as the training loop evaluates it
    (reference `src/train_lightgcn.py:141`, `src/lightgcn.py:279-286`): the mean of
    -logsigmoid(pos-neg), divided by n_pairs and multiplied back by the batch size."""
    n_pairs = pos_rank.size(0)
    log_prob = F.logsigmoid(pos_rank - neg_rank).mean()
    return (-log_prob + 0) / n_pairs * n_pairs


def regularization_loss(init_embed: Tensor, batch_size: int, users: Tensor, pos: Tensor,
                        neg: Tensor, decay: float) -> Tensor:
    """L2 term on the layer-0 rows, duplicates counted (reference `src/utils_v2.py:193-211`)."""
    sq = (init_embed[users].norm().pow(2) + init_embed[pos].norm().pow(2)
          + init_embed[neg].norm().pow(2))
    return (1 / 2) * sq / batch_size * decay


def train_step(model: PortLightGCN, optimizer: torch.optim.Optimizer, edge_index: Tensor,
               edge_weight: Tensor, users: Tensor, pos: Tensor, neg: Tensor,
               decay: float) -> Tuple[float, float, float]:
    """One iteration of `mini_batch_loop` (reference `src/train_lightgcn.py:129-151`) on
    pre-sampled triples. Returns (bpr, reg, total) like the three `.item()` reads."""
    optimizer.zero_grad()
    labels = batch_pos_neg_edges(users, pos, neg)
    out = model(edge_index, labels, edge_weight)
    size = len(users)
    bpr = bpr_loss(out[:size], out[size:])
    reg = regularization_loss(model.embedding.weight, size, users, pos, neg, decay)
    loss = bpr + reg
    loss.backward()
    optimizer.step()
    return bpr.item(), reg.item(), loss.item()


# --------------------------------------------------------------------------- scoring
def dense_seen_mask(seen_ptr: np.ndarray, seen_items: np.ndarray, n_items: int) -> Tensor:
    """Dense float mask [U, n_items] the reference feeds to `recommendK`
    (`interact_matrix` + `index_select(...).to_dense()`, `src/utils_v2.py:92-103,137-138`)."""
    n = len(seen_ptr) - 1
    mask = torch.zeros(n, n_items, dtype=torch.float32)
    rows = np.repeat(np.arange(n), np.diff(seen_ptr))
    mask[torch.from_numpy(rows), torch.from_numpy(np.asarray(seen_items, dtype=np.int64))] = 1.0
    return mask


def recommend_topk(embeds: Tensor, n_users: int, n_items: int, interactions_t: Tensor,
                   user_id_list: Sequence[int], k: int) -> Tensor:
    """`recommendK` after `get_embedding` (reference `src/lightgcn.py:172-177`): user x item
    scores, MULTIPLICATIVE seen-mask (a seen item scores 0.0, not -inf), top-k item indices."""
    src, dst = torch.split(embeds, [n_users, n_items])
    pred = src[list(user_id_list)] @ dst.t()
    masked = torch.mul(pred, (1 - interactions_t.to(pred.dtype)))
    return masked.topk(k, dim=-1).indices


def masked_scores(embeds: Tensor, n_users: int, n_items: int, interactions_t: Tensor,
                  user_id_list: Sequence[int]) -> Tensor:
    src, dst = torch.split(embeds, [n_users, n_items])
    pred = src[list(user_id_list)] @ dst.t()
    return torch.mul(pred, (1 - interactions_t.to(pred.dtype)))


def mark_mapk(heldout_lists: List[Sequence[int]], topk: np.ndarray, k: int) -> Tuple[float, float]:
    """Mean precision@k / recall@k with the set semantics of `MARK_MAPK`
    (reference `src/lightgcn.py:184-189`); row i of `topk` belongs to `heldout_lists[i]`."""
    prec, rec = [], []
    for held, top in zip(heldout_lists, topk):
        overlap = len(set(int(x) for x in held).intersection(int(x) for x in top))
        rec.append(overlap / len(held))
        prec.append(overlap / k)
    return float(np.mean(prec)), float(np.mean(rec))


# --------------------------------------------------------------------------- CSR (integer part)
def csr_by_target(edge_index: np.ndarray, edge_weight: np.ndarray, num_nodes: int):
    """The bit-exact targets for the graph-build kernels: destination-major CSR with the edge
    order of each row preserved (stable), the count degree, and the fp32 weighted degree /
    deg^-1/2 / w_hat that `gcn_norm` produces when it accumulates in edge order (CPU
    `scatter_add_` is serial along the scatter dimension)."""
    row = np.asarray(edge_index[0], dtype=np.int64)
    col = np.asarray(edge_index[1], dtype=np.int64)
    w = np.asarray(edge_weight, dtype=np.float32)
    order = np.argsort(col, kind="stable")
    counts = np.bincount(col, minlength=num_nodes).astype(np.int64)
    rowptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    deg = np.zeros(num_nodes, dtype=np.float32)
    np.add.at(deg, col, w)                                  # sequential, in edge order
    with np.errstate(divide="ignore"):
        dis = (np.float32(1.0) / np.sqrt(deg)).astype(np.float32)
    dis[np.isinf(dis)] = 0
    w_hat = (dis[row] * w) * dis[col]
    return {"rowptr": rowptr, "src": row[order], "eid": order, "count_deg": counts,
            "deg": deg, "dis": dis, "w_hat_edge": w_hat, "w_hat_csr": w_hat[order]}
