"""Device-side BPR sampler: `batch_loader` of the reference (`src/utils_v2.py:168-181`) as one
kernel launch on resident CSR purchase / ignore lists.

    sampler = DeviceSampler.from_lists(users, pos_ptr, pos_items, ign_ptr, ign_items, n_users, n_items, "cuda")
    u, p, n = sampler.sample(1024)            # int64 device tensors, same meaning as batch_loader's

`from_frame` accepts the reference's own `train_pos_list_df` (columns user_id_idx,
item_id_idx_list, ignor_neg_list) so `train_lightgcn.py` can switch with one line."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch
from torch import Tensor

from . import _capi
from .graph import _ptr, _stream


class DeviceSampler:
    def __init__(self, purchasers: Tensor, pos_ptr: Tensor, pos_items: Tensor, ign_ptr: Tensor, ign_items: Tensor,
                 n_users: int, n_items: int, seed: int = 0):
        for t in (purchasers, pos_ptr, pos_items, ign_ptr, ign_items):
            if not (t.is_cuda and t.dtype == torch.int64 and t.is_contiguous()):
                raise ValueError("DeviceSampler needs contiguous int64 CUDA tensors (no CPU fallback)")
        self.purchasers, self.pos_ptr, self.pos_items = purchasers, pos_ptr, pos_items
        self.ign_ptr, self.ign_items = ign_ptr, ign_items
        self.n_users, self.n_items, self.seed, self.step = int(n_users), int(n_items), int(seed), 0
        self._lib = _capi.lib()
        self._ws = None

    @staticmethod
    def from_lists(users: Sequence[int], pos_ptr, pos_items, ign_ptr, ign_items, n_users: int, n_items: int,
                   device, seed: int = 0) -> "DeviceSampler":
        ign_ptr = np.asarray(ign_ptr, dtype=np.int64)
        ign_items = np.asarray(ign_items, dtype=np.int64).copy()
        for i in range(len(ign_ptr) - 1):                       # rows must be sorted (binary search)
            seg = ign_items[ign_ptr[i]:ign_ptr[i + 1]]
            if seg.size > 1 and np.any(seg[1:] < seg[:-1]):
                seg.sort()
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.int64)), device=device)
        return DeviceSampler(dev(users), dev(pos_ptr), dev(pos_items), dev(ign_ptr), dev(ign_items), n_users, n_items,
                             seed)

    @staticmethod
    def from_frame(train_pos_list_df, n_users: int, n_items: int, device, seed: int = 0) -> "DeviceSampler":
        """From the reference's `train_pos_list_df` (`prepare_val_test`, src/utils_v2.py:139-141)."""
        users = train_pos_list_df["user_id_idx"].to_numpy(dtype=np.int64)
        pos_lists = [np.asarray(x, dtype=np.int64) for x in train_pos_list_df["item_id_idx_list"]]
        ign_lists = [np.sort(np.asarray(list(x), dtype=np.int64)) for x in train_pos_list_df["ignor_neg_list"]]
        ptr = lambda ls: np.concatenate([[0], np.cumsum([len(x) for x in ls])]).astype(np.int64)
        cat = lambda ls: np.concatenate(ls) if ls else np.zeros(0, np.int64)
        return DeviceSampler.from_lists(users, ptr(pos_lists), cat(pos_lists), ptr(ign_lists), cat(ign_lists),
                                        n_users, n_items, device, seed)

    def sample(self, batch: int):
        dev = self.purchasers.device
        out = torch.empty(3, batch, dtype=torch.int64, device=dev)
        need = self._lib.lgc_sample_triples_workspace_bytes(self.purchasers.numel(), batch)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = self._lib.lgc_sample_triples(self.purchasers.numel(), _ptr(self.purchasers), _ptr(self.pos_ptr),
                                              _ptr(self.pos_items), _ptr(self.ign_ptr), _ptr(self.ign_items),
                                              self.n_users, self.n_items, batch, self.seed, self.step, _ptr(out[0]),
                                              _ptr(out[1]), _ptr(out[2]), _ptr(self._ws), self._ws.numel(), _stream())
        if rc != 0:
            msg = self._lib.lgc_last_error().decode("utf-8", "replace")
            if "larger than population" in msg:
                raise ValueError(msg)                            # what random.sample raises
            raise RuntimeError(f"lgc_sample_triples failed with status {rc}: {msg}")
        self.step += 1
        return out[0], out[1], out[2]
