"""Serving adapter with the contract of the reference's TorchServe handler
(`LightGCNHandler.inference`, reference `torchserve/lightgcn_handler.py:73-96`):

    request  : a list of user indices          [user_idx, ...]
    response : {'items': [[20 item indices], ...]}   (item ids 0..n_items-1, best first)

What changes underneath: the final embeddings are propagated ONCE per model
version (`LightGCN.cached_embedding`) instead of on every request (`recommendK` calls
`get_embedding`, `:91`), and the dense `[len(request), n_items]` seen-mask the handler builds with
`index_select(...).to_dense().cpu()` (`:88`) is a CSR of the train purchases that stays on the device.
The reference's own handler class also runs unchanged over the drop-in `LightGCN` (tests/test_gpu_dropin.py).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch
from torch import Tensor

from . import scoring


class LightGCNService:
    def __init__(self, model, edge_index: Tensor, edge_weight: Tensor, n_users: int, n_items: int,
                 seen_ptr: Tensor, seen_items: Tensor, k: int = 20):
        """`seen_ptr` [n_users + 1] / `seen_items`: CSR over ALL users of their train purchases
        (weight == 1.0 edges, un-offset item ids) -- the sparse form of the handler's `i_m_matrix`."""
        if not edge_index.is_cuda:
            raise RuntimeError("LightGCNService needs the graph on a CUDA device (no CPU fallback)")
        self.model, self.edge_index, self.edge_weight = model, edge_index, edge_weight
        self.n_users, self.n_items, self.k = int(n_users), int(n_items), int(k)
        dev = edge_index.device
        self.seen_ptr = seen_ptr.to(device=dev, dtype=torch.int64).contiguous()
        self.seen_items = seen_items.to(device=dev, dtype=torch.int64).contiguous()

    @staticmethod
    def from_train_frame(model, train_df, device, k: int = 20, handler_graph_quirk: bool = False) -> "LightGCNService":
        """From the `processed_train.csv` frame the handler loads (`:33-41`): columns user_id_idx,
        item_id_idx (offset by n_users), weight.

        `handler_graph_quirk`: the reference handler subtracts n_users from the item column BEFORE it
        builds the graph (`:38-40`), so its edge list uses un-offset item ids that alias the first
        n_items users -- unlike training (`src/train_lightgcn.py:35`) and `InferenceLightGCN`
        (`src/inference_lightgcn.py:18`), which build the graph on offset ids. Default: the graph the
        model was trained on; True reproduces the handler's lists exactly."""
        n_users = int(train_df['user_id_idx'].nunique())
        n_items = int(train_df['item_id_idx'].nunique())
        u = torch.as_tensor(train_df['user_id_idx'].to_numpy(dtype=np.int64))
        i = torch.as_tensor(train_df['item_id_idx'].to_numpy(dtype=np.int64))
        w = torch.as_tensor(train_df['weight'].to_numpy(dtype=np.float32))
        gi = i - n_users if handler_graph_quirk else i
        edge_index = torch.stack((torch.cat([u, gi]), torch.cat([gi, u]))).to(device)   # df_to_graph, :112-131
        edge_weight = torch.cat([w, w]).to(device)
        bought = w == 1.0                                                             # interact_matrix, :133-144
        key = torch.unique(u[bought] * n_items + (i[bought] - n_users))
        ku, ki = key // n_items, key % n_items
        ptr = torch.zeros(n_users + 1, dtype=torch.int64)
        ptr[1:] = torch.cumsum(torch.bincount(ku, minlength=n_users), 0)
        return LightGCNService(model, edge_index, edge_weight, n_users, n_items, ptr, ki, k)

    def _seen_of(self, users: Tensor) -> scoring.SeenLists:
        lo, hi = self.seen_ptr[users], self.seen_ptr[users + 1]
        cnt = hi - lo
        ptr = torch.zeros(users.numel() + 1, dtype=torch.int64, device=users.device)
        ptr[1:] = torch.cumsum(cnt, 0)
        idx = torch.repeat_interleave(lo - ptr[:-1], cnt) + torch.arange(int(ptr[-1]), device=users.device)
        return scoring.SeenLists(ptr, self.seen_items[idx])

    def inference(self, data: Sequence[int]) -> Dict[str, List[List[int]]]:
        """`LightGCNHandler.inference`: [user_idx, ...] -> {'items': [[k item idx], ...]}."""
        with torch.no_grad():
            users = torch.as_tensor(np.asarray(data, dtype=np.int64), device=self.edge_index.device)
            if users.numel() and (int(users.min()) < 0 or int(users.max()) >= self.n_users):
                raise IndexError("user index out of range")          # what index_select raises in the reference
            top = self.model.recommendK_array(self.edge_index, self.edge_weight, self.n_users, self.n_items,
                                              self._seen_of(users), users, self.k)
        return {'items': top.tolist()}

    def handle(self, data, context=None):
        """preprocess -> inference -> postprocess of the handler (`:54-110`)."""
        body = data[0].get("data")
        if body is None:
            body = data[0].get("body")
        return [self.inference(body)]
