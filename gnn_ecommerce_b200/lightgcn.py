"""Drop-in `LightGCN` / `LGConv` / `BPRLoss` for the reference's module seam.

Same constructor, attributes, method names, argument meaning and `state_dict()` keys
(`alpha`, `embedding.weight`) as reference `src/lightgcn.py:13-286`, so `train_lightgcn.py`,
`inference_lightgcn.py` and the TorchServe handler can import this module instead. What changes
is underneath: `get_embedding` is one fused K-layer CSR SpMM chain on sm_100a (graph normalised
once, not K times per call), its backward is the same chain on the gradient, and `recommendK`
is a tcgen05 GEMM with a fused candidate filter + exact fp32 re-scoring instead of a dense
[U, I] score matrix copied to the host. CUDA only: CPU tensors raise (there is no fallback).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np
import pandas as pd
import torch
import torch.nn.functional as F
from torch import Tensor
from torch.nn import Embedding, ModuleList

from . import ops, scoring
from .graph import Graph, graph_for

Adj = Tensor
OptTensor = Optional[Tensor]


def _require_cuda(t: Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} lives on {t.device}: gnn_ecommerce_b200 has no CPU path "
                           "(move the model and the graph to a CUDA device)")


class LGConv(torch.nn.Module):
    """Operator seam: `LGConv(normalize=True).forward(x, edge_index, edge_weight=None)`
    (PyG operator the reference builds at `src/lightgcn.py:82` and calls at `:96`).
    x'_i = sum_{j->i} w_ji / sqrt(deg_i deg_j) x_j with the weighted in-degree; no parameters;
    differentiable w.r.t. x."""

    def __init__(self, normalize: bool = True, **kwargs):
        super().__init__()
        self.normalize = normalize

    def reset_parameters(self):
        pass

    def forward(self, x: Tensor, edge_index: Adj, edge_weight: OptTensor = None) -> Tensor:
        _require_cuda(x, "x")
        _require_cuda(edge_index, "edge_index")
        g = graph_for(edge_index, edge_weight, x.size(0), self.normalize)
        return ops.spmm_autograd(x, g)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"


class LightGCN(torch.nn.Module):
    """LightGCN with the reference's interface (`src/lightgcn.py:58-231`).

    Args mirror the reference: `num_nodes`, `embedding_dim`, `num_layers`, optional `alpha`
    (float or tensor of K+1 layer weights, default uniform 1/(K+1)), `**kwargs` forwarded to the
    `LGConv` layers.
    """

    def __init__(self, num_nodes: int, embedding_dim: int, num_layers: int,
                 alpha: Optional[Union[float, Tensor]] = None, **kwargs):
        super().__init__()
        self.num_nodes = num_nodes
        self.embedding_dim = embedding_dim
        self.num_layers = num_layers
        if alpha is None:
            alpha = 1. / (num_layers + 1)
        if isinstance(alpha, Tensor):
            assert alpha.size(0) == num_layers + 1
        else:
            alpha = torch.tensor([alpha] * (num_layers + 1))
        self.register_buffer('alpha', alpha)
        self.embedding = Embedding(num_nodes, embedding_dim)
        self.convs = ModuleList([LGConv(**kwargs) for _ in range(num_layers)])
        self._normalize = bool(kwargs.get("normalize", True))
        self._alpha_host = None
        self._emb_cache = None          # (key, table): final embeddings for serving / evaluation
        self._weights_epoch = 0         # bumped by FusedBPRTrainer (it updates the table through raw pointers)
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.xavier_uniform_(self.embedding.weight)
        for conv in self.convs:
            conv.reset_parameters()

    # ------------------------------------------------------------------ helpers
    def alpha_host(self) -> List[float]:
        """Layer weights as host floats (cached; one device read per change of `alpha`)."""
        key = (self.alpha.data_ptr(), self.alpha._version)
        if self._alpha_host is None or self._alpha_host[0] != key:
            self._alpha_host = (key, [float(a) for a in self.alpha.detach().float().cpu().tolist()])
        return self._alpha_host[1]

    def graph(self, edge_index: Adj, edge_weight: OptTensor) -> Graph:
        _require_cuda(edge_index, "edge_index")
        return graph_for(edge_index, edge_weight, self.num_nodes, self._normalize)

    def cached_embedding(self, edge_index: Adj, edge_weight: OptTensor) -> Tensor:
        """Final embeddings for scoring, computed once per (weights, graph) instead of on every
        `recommendK` / TorchServe request as the reference does (`src/lightgcn.py:171`,
        `torchserve/lightgcn_handler.py:91`). No autograd graph is kept."""
        w = self.embedding.weight
        key = (w.data_ptr(), w._version, self._weights_epoch, edge_index.data_ptr(), edge_index._version,
               tuple(edge_index.shape), None if edge_weight is None else (edge_weight.data_ptr(), edge_weight._version),
               self.alpha._version, self.num_layers)
        if self._emb_cache is None or self._emb_cache[0] != key:
            with torch.no_grad():
                self._emb_cache = (key, self.get_embedding(edge_index, edge_weight).detach())
        return self._emb_cache[1]

    # ------------------------------------------------------------------ reference API
    def get_embedding(self, edge_index: Adj, edge_weight: OptTensor) -> Tensor:
        """sum_l alpha_l A_hat^l E0 (reference `src/lightgcn.py:91-99`), differentiable."""
        _require_cuda(self.embedding.weight, "embedding.weight")
        g = self.graph(edge_index, edge_weight)
        return ops.propagate_autograd(self.embedding.weight, g, self.alpha_host(), self.num_layers)

    def forward(self, edge_index: Adj, edge_label_index: OptTensor = None,
                edge_weight: OptTensor = None) -> Tensor:
        """Rankings <out[a], out[b]> for the node pairs in `edge_label_index`
        (reference `src/lightgcn.py:101-125`; default: the graph's own edges)."""
        if edge_label_index is None:
            edge_label_index = edge_index
        out = self.get_embedding(edge_index, edge_weight)
        return (out[edge_label_index[0]] * out[edge_label_index[1]]).sum(dim=-1)

    def predict_link(self, edge_index: Adj, edge_label_index: OptTensor = None,
                     prob: bool = False) -> Tensor:
        pred = self(edge_index, edge_label_index).sigmoid()
        return pred if prob else pred.round()

    def recommend(self, edge_index: Adj, src_index: OptTensor = None, dst_index: OptTensor = None,
                  k: int = 1, edge_weight: OptTensor = None) -> Tensor:
        """Top-k of out[src] @ out[dst]^T (reference `src/lightgcn.py:138-167`; the reference
        forgets `edge_weight` there and raises -- it is an optional extra argument here)."""
        out = self.cached_embedding(edge_index, edge_weight)
        src = torch.arange(self.num_nodes, device=out.device) if src_index is None else src_index
        dst = out if dst_index is None else out[dst_index]
        items, _ = scoring.score_topk(out, dst.contiguous(), src, None, None, k)
        if dst_index is not None:
            items = dst_index[items.view(-1)].view(*items.size())
        return items

    def recommendK_array(self, edge_index, edge_weight, n_users, n_items, interactions_t, user_id_list,
                         k: int = 5, to_host: bool = True):
        """`recommendK` without the pandas packaging: the top-k item ids as one `[len(user_id_list), k]`
        array -- a host ndarray (int32, through pinned memory) or, with `to_host=False`, the int64 device
        tensor. `user_id_list` may be a list, an ndarray or a tensor (pinned host tensors are copied
        asynchronously)."""
        embeds = self.cached_embedding(edge_index, edge_weight)
        if isinstance(user_id_list, Tensor):
            users = user_id_list.to(device=embeds.device, dtype=torch.int64, non_blocking=True)
        else:
            users = torch.as_tensor(np.asarray(user_id_list, dtype=np.int64), device=embeds.device)
        seen = scoring.as_seen_lists(interactions_t, users.numel(), n_items, embeds.device)
        rows = ops.full_rows(embeds)
        items, _ = scoring.score_topk(rows[:n_users], rows[n_users:n_users + n_items], users,
                                      seen.ptr, seen.items, k, d=self.embedding_dim)
        return scoring.topk_to_host(items) if to_host else items

    def recommendK(self, edge_index, edge_weight, n_users, n_items, interactions_t, user_id_list,
                   k: int = 5):
        """Top-k unseen-first item lists for `user_id_list` (reference `src/lightgcn.py:169-182`).

        `interactions_t` is the reference's dense 0/1 mask [len(user_id_list), n_items] (any
        device) or a `scoring.SeenLists` CSR; masking is multiplicative like the reference's
        (`pred * (1 - mask)`: a seen item scores 0.0). Returns the same two-column frame
        (`user_ID`, `top_rlvnt_itm`)."""
        top = self.recommendK_array(edge_index, edge_weight, n_users, n_items, interactions_t, user_id_list, k)
        top_index_df = pd.DataFrame({'user_ID': list(user_id_list), 'top_rlvnt_itm': top.tolist()})
        return top_index_df[['user_ID', 'top_rlvnt_itm']]

    def MARK_MAPK(self, test_pos_list_df, top_index_df, k):
        """Mean precision@k / recall@k over the users of `test_pos_list_df`, plus the per-user
        frame, with the set semantics of reference `src/lightgcn.py:184-189`."""
        metrics = pd.merge(test_pos_list_df, top_index_df, how='left', left_on='user_id_idx',
                           right_on='user_ID')
        held, top = metrics['item_id_idx_list'].tolist(), metrics['top_rlvnt_itm'].tolist()
        overlap = [list(set(h).intersection(t)) for h, t in zip(held, top)]
        metrics['overlap_item'] = overlap
        metrics['recall'] = [len(o) / len(h) for o, h in zip(overlap, held)]
        metrics['precision'] = [len(o) / k for o in overlap]
        return metrics['precision'].mean(), metrics['recall'].mean(), metrics

    def link_pred_loss(self, pred: Tensor, edge_label: Tensor, **kwargs) -> Tensor:
        return torch.nn.BCEWithLogitsLoss(**kwargs)(pred, edge_label.to(pred.dtype))

    def recommendation_loss(self, pos_edge_rank: Tensor, neg_edge_rank: Tensor,
                            lambda_reg: float = 1e-4, **kwargs) -> Tensor:
        return BPRLoss(lambda_reg, **kwargs)(pos_edge_rank, neg_edge_rank, self.embedding.weight)

    def __repr__(self) -> str:
        return (f'{self.__class__.__name__}({self.num_nodes}, '
                f'{self.embedding_dim}, num_layers={self.num_layers})')


class BPRLoss(torch.nn.modules.loss._Loss):
    """Mean BPR loss with the reference's scaling (`src/lightgcn.py:234-286`):
    (-mean(logsigmoid(pos - neg)) + lambda * ||parameters||^2) / n_pairs."""
    __constants__ = ['lambda_reg']
    lambda_reg: float

    def __init__(self, lambda_reg: float = 0, **kwargs) -> None:
        super().__init__(None, None, "sum", **kwargs)
        self.lambda_reg = lambda_reg

    def forward(self, positives: Tensor, negatives: Tensor, parameters: Tensor = None) -> Tensor:
        n_pairs = positives.size(0)
        nll = -F.logsigmoid(positives - negatives).mean()
        reg = 0 if self.lambda_reg == 0 else self.lambda_reg * parameters.norm(p=2).pow(2)
        return (nll + reg) / n_pairs
