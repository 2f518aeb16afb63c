"""Top-k scoring wrapper (`recommendK`, reference `src/lightgcn.py:169-182`).

Host side only converts the reference's dense seen-mask into CSR seen-lists and hands raw
pointers to `lgc_score_topk`; the GEMM, the candidate filter, the exact fp32 re-scoring and the
multiplicative mask all run in `csrc/score.cu`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _capi
from .graph import _ptr, _stream


@dataclass
class SeenLists:
    """CSR of already-seen (train purchase) items per scored user: row i belongs to
    `user_id_list[i]`. Sparse form of `interactions_t` (reference `src/utils_v2.py:92-103,137`)."""
    ptr: Optional[Tensor]      # int64 [U+1] on the device, or None (nothing seen)
    items: Optional[Tensor]    # int64, un-offset item ids

    @staticmethod
    def from_numpy(ptr: np.ndarray, items: np.ndarray, device) -> "SeenLists":
        return SeenLists(torch.as_tensor(np.asarray(ptr, dtype=np.int64), device=device),
                         torch.as_tensor(np.asarray(items, dtype=np.int64), device=device))


def as_seen_lists(interactions_t, n_rows: int, n_items: int, device) -> SeenLists:
    if interactions_t is None:
        return SeenLists(None, None)
    if isinstance(interactions_t, SeenLists):
        return interactions_t
    t = interactions_t
    if t.is_sparse:
        t = t.coalesce()
        idx, val = t.indices(), t.values()
        if not bool(((val == 0) | (val == 1)).all()):
            raise NotImplementedError("recommendK: the seen-mask must be 0/1")
        keep = val != 0
        rows, cols = idx[0][keep], idx[1][keep]
    else:
        if tuple(t.shape) != (n_rows, n_items):
            raise ValueError(f"interactions_t must be [{n_rows}, {n_items}], got {tuple(t.shape)}")
        if not bool(((t == 0) | (t == 1)).all()):
            raise NotImplementedError("recommendK: the seen-mask must be 0/1")
        nz = t.nonzero()
        rows, cols = nz[:, 0], nz[:, 1]            # row-major order: rows ascend
    counts = torch.bincount(rows, minlength=n_rows)
    ptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=rows.device)
    ptr[1:] = torch.cumsum(counts, 0)
    order = torch.argsort(rows, stable=True)
    return SeenLists(ptr.to(device), cols[order].to(device=device, dtype=torch.int64).contiguous())


def _aligned_rows(t: Tensor, d: int) -> Tensor:
    """Rows must start on 16-byte boundaries (128-bit loads); re-pitch a table that does not
    (e.g. a contiguous [N, 90] tensor) into a zero-padded [N, ld] copy."""
    if t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0:
        return t
    ld = (d + 3) // 4 * 4
    out = torch.zeros(t.size(0), ld, dtype=torch.float32, device=t.device)
    out[:, :d] = t[:, :d]
    return out


_WORKSPACES = {}


def _workspace(nbytes: int, dev) -> Tensor:
    """The scoring workspace (bound records and candidate lists: ~16 GB for 1.6 M users) is kept per
    device and reused by later calls that fit: allocating and freeing it on every call made one call
    in three take 2.5x longer (the caching allocator returning the block to the driver)."""
    key = (dev.type, dev.index)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        _WORKSPACES.pop(key, None)
        ws = None
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _WORKSPACES[key] = ws
    return ws


def release_workspaces() -> None:
    """Free the cached scoring workspaces."""
    _WORKSPACES.clear()


def score_topk(user_emb: Tensor, item_emb: Tensor, user_ids: Optional[Tensor],
               seen_ptr: Optional[Tensor], seen_items: Optional[Tensor], k: int,
               d: Optional[int] = None, return_stats: bool = False, out: Optional[Tuple[Tensor, Tensor]] = None):
    """Top-k items per user by masked fp32 score. `user_emb` / `item_emb` are row tables
    (row stride = `stride(0)` floats, multiple of 4); `d` = number of leading columns used.
    `out` = (items int64 [U, k], scores float32 [U, k]) of an earlier call to write into: a caller that
    scores the same number of users repeatedly (serving, evaluation every epoch) then allocates nothing --
    a fresh 384 MB result for 1.6 M users is a cudaMalloc inside the call whenever the previous result is
    still referenced (measured: one call in three 1.2x-9x slower)."""
    lib = _capi.lib()
    for t, name in ((user_emb, "user_emb"), (item_emb, "item_emb")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1):
            raise ValueError(f"{name} must be a 2-D float32 CUDA row table (no CPU fallback)")
    d = int(d if d is not None else min(user_emb.size(1), item_emb.size(1)))
    user_emb, item_emb = _aligned_rows(user_emb, d), _aligned_rows(item_emb, d)
    n_items = item_emb.size(0)
    if user_ids is None:
        n_users = user_emb.size(0)
    else:
        user_ids = user_ids.to(device=user_emb.device, dtype=torch.int64).contiguous()
        n_users = user_ids.numel()
    if not 1 <= k <= min(32, n_items):
        raise ValueError("k must be in 1..min(32, n_items)")
    dev = user_emb.device
    if out is not None:
        items, scores = out
        for t, dt in ((items, torch.int64), (scores, torch.float32)):
            if not (t.is_cuda and t.device == dev and t.dtype == dt and tuple(t.shape) == (n_users, k) and t.is_contiguous()):
                raise ValueError("out must be (int64 [U, k], float32 [U, k]) contiguous tensors on the tables' device")
    else:
        items = torch.empty(n_users, k, dtype=torch.int64, device=dev)
        scores = torch.empty(n_users, k, dtype=torch.float32, device=dev)
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    if n_users == 0:
        return (items, scores, stats) if return_stats else (items, scores)
    ws = _workspace(lib.lgc_score_topk_workspace_bytes(n_users, n_items, d, k), dev)
    args = _capi.ScoreTopkArgs(
        d=d, ld_user=user_emb.stride(0), ld_item=item_emb.stride(0), k=k, n_users=n_users,
        n_items=n_items, user_emb=_ptr(user_emb), item_emb=_ptr(item_emb), user_ids=_ptr(user_ids),
        seen_ptr=_ptr(seen_ptr), seen_items=_ptr(seen_items), topk_items=_ptr(items),
        topk_scores=_ptr(scores), stats=_ptr(stats), workspace=_ptr(ws), workspace_bytes=ws.numel())
    with torch.cuda.device(dev):
        rc = lib.lgc_score_topk(C.byref(args), _stream())
    _capi.check(rc, "lgc_score_topk")
    return (items, scores, stats) if return_stats else (items, scores)


_PINNED = {}


def topk_to_host(items: Tensor, dtype=torch.int32) -> np.ndarray:
    """Top-k item ids [U, k] -> host ndarray through a cached pinned buffer (one asynchronous D2H copy
    + one synchronisation). int32 by default: item ids fit, and the copy is what an all-user call is
    bound by on the host side (1.6 M x 20 ids = 128 MB). The reference does
    `.cpu().numpy().tolist()` on a [U, n_items] score matrix instead (src/lightgcn.py:170-177)."""
    dev_items = items.to(dtype) if items.dtype != dtype else items
    key = (dtype, items.device.index)
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < dev_items.numel():
        buf = torch.empty(max(dev_items.numel(), 1), dtype=dtype).pin_memory()
        _PINNED[key] = buf
    host = buf[:dev_items.numel()].view(dev_items.shape)
    host.copy_(dev_items, non_blocking=True)
    torch.cuda.current_stream(items.device).synchronize()
    return host.numpy()


def mark_mapk(topk_items: Tensor, held_ptr: Tensor, held_items: Tensor):
    """Device-side `MARK_MAPK` (reference `src/lightgcn.py:184-189`): returns
    `(mean precision@k, mean recall@k, per_user [U, 2])` for top-k lists already on the device
    (row i of `topk_items` and row i of the held-out CSR belong to the same user)."""
    lib = _capi.lib()
    dev = topk_items.device
    if not topk_items.is_cuda:
        raise RuntimeError("mark_mapk needs CUDA tensors (no CPU fallback)")
    top = topk_items.to(torch.int64).contiguous()
    ptr = held_ptr.to(device=dev, dtype=torch.int64).contiguous()
    items = held_items.to(device=dev, dtype=torch.int64).contiguous()
    n, k = top.shape
    per_user = torch.empty(n, 2, dtype=torch.float32, device=dev)
    out2 = torch.empty(2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.lgc_mark_mapk(n, k, _ptr(top), _ptr(ptr), _ptr(items), _ptr(per_user), _ptr(out2), _stream())
    _capi.check(rc, "lgc_mark_mapk")
    p, r = out2.cpu().tolist()
    return p, r, per_user
