"""Interaction-frame ingest on the device: the relabelling and the seen-list build
of the reference's `prepare_val_test` (reference `src/utils_v2.py:40-61,92-103,106-143`) without pandas
group-bys or a dense `[U, n_items]` mask, straight into what the kernels consume.

    ids = relabel(raw_user, raw_item)                    # == LabelEncoder.fit_transform on both columns
    graph = Graph.from_interactions(ids.user_idx, ids.item_idx + ids.n_users, weight, ids.n_users + ids.n_items)
    seen = seen_lists(ids.user_idx, ids.item_idx, weight, ids.n_users, ids.n_items)      # CSR of interact_matrix

Numbering is bit-identical to the reference's: `sklearn.preprocessing.LabelEncoder` numbers the
distinct raw ids in ascending order (`np.unique`), which is what a sorted `torch.unique` returns.
Everything here is index plumbing on torch tensors (sort / unique / bincount on the tensors' own
device); the arithmetic stays in the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from .scoring import SeenLists


@dataclass
class Relabelled:
    user_idx: Tensor          # int64 [E]  0..n_users-1       (`user_id_idx`)
    item_idx: Tensor          # int64 [E]  0..n_items-1       (`item_id_idx` BEFORE the + n_users offset)
    user_classes: Tensor      # sorted distinct raw user ids  (LabelEncoder.classes_)
    item_classes: Tensor
    n_users: int
    n_items: int

    def transform_users(self, raw: Tensor) -> Tensor:
        """`le_user.transform(raw)` (`src/utils_v2.py:53-57`): raises on an id the train split never saw."""
        return _transform(self.user_classes, raw, "user")

    def transform_items(self, raw: Tensor) -> Tensor:
        return _transform(self.item_classes, raw, "item")


def _transform(classes: Tensor, raw: Tensor, what: str) -> Tensor:
    raw = raw.to(classes.device)
    pos = torch.searchsorted(classes, raw)
    ok = (pos < classes.numel()) & (classes[pos.clamp_max(classes.numel() - 1)] == raw)
    if not bool(ok.all()):
        raise ValueError(f"y contains previously unseen {what} labels")      # LabelEncoder's own error
    return pos


def relabel(raw_user: Tensor, raw_item: Tensor) -> Relabelled:
    """`relabelling(train_df)` (`src/utils_v2.py:40-61`) for integer raw ids: dense ids in ascending raw-id
    order, per column."""
    uc, ui = torch.unique(raw_user, sorted=True, return_inverse=True)
    ic, ii = torch.unique(raw_item, sorted=True, return_inverse=True)
    return Relabelled(ui, ii, uc, ic, int(uc.numel()), int(ic.numel()))


def seen_lists(user_idx: Tensor, item_idx: Tensor, weight: Tensor, n_users: int, n_items: int,
               users: Optional[Tensor] = None) -> SeenLists:
    """CSR form of `interact_matrix` (`src/utils_v2.py:92-103`: the train edges with weight == 1.0) for
    `users` (default: every user), i.e. of the dense rows the reference builds with
    `index_select(...).to_dense()` (`:137-138`). Rows hold sorted distinct un-offset item ids."""
    bought = weight == 1.0
    key = torch.unique(user_idx[bought] * n_items + item_idx[bought])        # sorted, duplicates merged
    ku, ki = key // n_items, key % n_items
    if users is None:
        ptr = torch.zeros(n_users + 1, dtype=torch.int64, device=key.device)
        ptr[1:] = torch.cumsum(torch.bincount(ku, minlength=n_users), 0)
        return SeenLists(ptr, ki)
    users = users.to(device=key.device, dtype=torch.int64)
    lo = torch.searchsorted(ku, users, right=False)
    hi = torch.searchsorted(ku, users, right=True)
    cnt = hi - lo
    ptr = torch.zeros(users.numel() + 1, dtype=torch.int64, device=key.device)
    ptr[1:] = torch.cumsum(cnt, 0)
    idx = torch.repeat_interleave(lo - ptr[:-1], cnt) + torch.arange(int(ptr[-1]), device=key.device)
    return SeenLists(ptr, ki[idx])


def positive_lists(user_idx: Tensor, item_idx: Tensor, weight: Tensor, n_users: int):
    """`pos_item_list` (`src/utils_v2.py:64-73`) as CSR: (users with >= 1 weight == 1 row, ptr, items in
    frame order) -- the held-out lists `MARK_MAPK` and the device sampler consume."""
    bought = weight == 1.0
    u, it = user_idx[bought], item_idx[bought]
    order = torch.argsort(u, stable=True)
    u, it = u[order], it[order]
    users, counts = torch.unique_consecutive(u, return_counts=True)
    ptr = torch.zeros(users.numel() + 1, dtype=torch.int64, device=u.device)
    ptr[1:] = torch.cumsum(counts, 0)
    return users, ptr, it
