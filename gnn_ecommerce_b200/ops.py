"""Thin torch-facing wrappers over the C ABI: tensors in, raw pointers + stream out.

Everything here is plumbing (device memory, streams, autograd glue); the arithmetic lives in
`csrc/*.cu`. All tensors must be CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Sequence

import torch
from torch import Tensor

from . import _capi
from .graph import Graph, _ptr, _stream, padded_dim


def _alpha_array(alpha: Sequence[float]):
    arr = (C.c_float * len(alpha))(*[float(a) for a in alpha])
    return arr


def _check_table(x: Tensor, g: Graph, name: str) -> int:
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2:
        raise ValueError(f"{name} must be a 2-D float32 CUDA tensor")
    if x.size(0) != g.num_nodes:
        raise ValueError(f"{name} has {x.size(0)} rows, the graph has {g.num_nodes} nodes")
    if x.stride(1) != 1 or x.stride(0) % 4 or x.data_ptr() % 16:
        raise ValueError(f"{name} rows must be 16-byte aligned runs of floats")
    ld = x.stride(0)
    if x.size(1) > ld or not _capi.lib().lgc_ld_supported(ld):
        raise ValueError(f"{name}: unsupported row stride {ld}")
    return ld


def pad_table(x: Tensor) -> Tensor:
    """Return `x` itself when its rows already form a supported padded table, else a zero-padded
    copy [N, ld]. Padding columns stay zero through every kernel (all ops are linear in them)."""
    ld = padded_dim(x.size(1))
    if (x.dim() == 2 and x.stride(1) == 1 and x.stride(0) == ld and x.data_ptr() % 16 == 0
            and (ld == x.size(1) or _pad_is_zero_view(x, ld))):
        return x
    out = new_table(x.size(0), x.size(1), x.device)     # [N, d] view of a zeroed [N, ld] table
    out.copy_(x)
    return out


# storage pointer -> weak reference to the zero-initialised base tensor `new_table` handed out. A
# view keeps its base alive (`._base`), so the entry lives exactly as long as some view of the
# table does; once the table is freed the entry disappears with it and a later allocation that
# reuses the address is NOT mistaken for a zero-padded table.
_PADDED_TABLES = {}


def _pad_is_zero_view(x: Tensor, ld: int) -> bool:
    ptr = x.untyped_storage().data_ptr()
    ref = _PADDED_TABLES.get(ptr)
    base = ref() if ref is not None else None
    return base is not None and base.untyped_storage().data_ptr() == ptr and base.stride(0) == ld


def new_table(n_rows: int, d: int, device) -> Tensor:
    """Zeroed [n_rows, d] view of a padded [n_rows, ld] allocation."""
    ld = padded_dim(d)
    store = torch.zeros(n_rows, ld, dtype=torch.float32, device=device)
    if ld == d:
        return store
    ptr = store.untyped_storage().data_ptr()
    _PADDED_TABLES[ptr] = weakref.ref(store, lambda _r, _p=ptr: _PADDED_TABLES.pop(_p, None))
    return store[:, :d]


def full_rows(x: Tensor) -> Tensor:
    """The [N, ld] table behind a padded [N, d] view."""
    ld = x.stride(0)
    return x if ld == x.size(1) else x.as_strided((x.size(0), ld), (ld, 1))


def spmm(g: Graph, x: Tensor) -> Tensor:
    """y = A_hat x for a padded table x [N, ld-strided]."""
    lib = _capi.lib()
    ld = _check_table(x, g, "x")
    y = new_table(g.num_nodes, x.size(1), x.device)
    ws_bytes = lib.lgc_spmm_workspace_bytes(g.handle, ld)
    ws = g.workspace(("spmm", ld), ws_bytes)
    with torch.cuda.device(x.device):
        rc = lib.lgc_spmm(g.handle, ld, _ptr(x), _ptr(y), _ptr(ws), ws.numel(), _stream())
    _capi.check(rc, "lgc_spmm")
    return y


def propagate(g: Graph, x0: Tensor, alpha: Sequence[float], num_layers: int) -> Tensor:
    """out = sum_l alpha_l A_hat^l x0 (reference `get_embedding`, src/lightgcn.py:91-99)."""
    lib = _capi.lib()
    ld = _check_table(x0, g, "x0")
    out = new_table(g.num_nodes, x0.size(1), x0.device)
    ws_bytes = lib.lgc_propagate_workspace_bytes(g.handle, ld, num_layers)
    ws = g.workspace(("propagate", ld, num_layers), ws_bytes)
    with torch.cuda.device(x0.device):
        rc = lib.lgc_propagate(g.handle, ld, num_layers, _alpha_array(alpha), _ptr(x0), _ptr(out),
                               _ptr(ws), ws.numel(), _stream())
    _capi.check(rc, "lgc_propagate")
    return out


class _SpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, g: Graph) -> Tensor:
        ctx.g, ctx.d = g, x.size(1)
        return spmm(g, pad_table(x.detach()))

    @staticmethod
    def backward(ctx, grad: Tensor):
        gx = spmm(ctx.g.transpose(), pad_table(grad))
        return gx, None


class _PropagateFn(torch.autograd.Function):
    """Differentiable `get_embedding`: the operator sum_l alpha_l A_hat^l is symmetric when A_hat
    is, so the backward pass is the same kernel chain on the incoming gradient."""

    @staticmethod
    def forward(ctx, x0: Tensor, g: Graph, alpha: tuple, num_layers: int) -> Tensor:
        ctx.g, ctx.alpha, ctx.k = g, alpha, num_layers
        return propagate(g, pad_table(x0.detach()), alpha, num_layers)

    @staticmethod
    def backward(ctx, grad: Tensor):
        gx = propagate(ctx.g.transpose(), pad_table(grad), ctx.alpha, ctx.k)
        return gx, None, None, None


def spmm_autograd(x: Tensor, g: Graph) -> Tensor:
    return _SpmmFn.apply(x, g)


def propagate_autograd(x0: Tensor, g: Graph, alpha: Sequence[float], num_layers: int) -> Tensor:
    return _PropagateFn.apply(x0, g, tuple(float(a) for a in alpha), num_layers)


def pair_scores(out: Tensor, pairs: Tensor) -> Tensor:
    lib = _capi.lib()
    out = pad_table(out)
    pairs = pairs.to(torch.int64).contiguous()
    score = torch.empty(pairs.size(1), dtype=torch.float32, device=out.device)
    with torch.cuda.device(out.device):
        rc = lib.lgc_pair_scores(out.stride(0), _ptr(out), _ptr(pairs), pairs.size(1), _ptr(score),
                                 _stream())
    _capi.check(rc, "lgc_pair_scores")
    return score


def adam_step(p: Tensor, grad: Tensor, m: Tensor, v: Tensor, lr: float, betas=(0.9, 0.999),
              eps: float = 1e-8, step: int = 1) -> None:
    """Single-pass dense torch.optim.Adam update on contiguous fp32 buffers (in place)."""
    lib = _capi.lib()
    for t in (p, grad, m, v):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("adam_step needs contiguous float32 CUDA tensors")
    with torch.cuda.device(p.device):
        rc = lib.lgc_adam_step(p.numel(), _ptr(p), _ptr(grad), _ptr(m), _ptr(v), lr, betas[0], betas[1],
                               eps, step, _stream())
    _capi.check(rc, "lgc_adam_step")


def bpr_loss_grad(out: Tensor, e0: Tensor, users: Tensor, pos: Tensor, neg: Tensor, decay: float,
                  alpha0: float, grad_out: Tensor, grad_e0: Optional[Tensor] = None):
    """loss3 = [bpr, reg, total]; accumulates dL/d out into grad_out and (optionally)
    alpha0 * dL/d out + L2 gradient into grad_e0. Tables are padded [N, ld]."""
    lib = _capi.lib()
    ld = out.stride(0)
    batch = users.numel()
    loss3 = torch.empty(3, dtype=torch.float32, device=out.device)
    ws = torch.empty(lib.lgc_bpr_workspace_bytes(batch), dtype=torch.uint8, device=out.device)
    with torch.cuda.device(out.device):
        rc = lib.lgc_bpr_loss_grad(out.size(0), ld, batch, _ptr(users), _ptr(pos), _ptr(neg), _ptr(out),
                                   _ptr(e0), float(decay), float(alpha0), _ptr(grad_out),
                                   _ptr(grad_e0), None, _ptr(loss3), _ptr(ws), ws.numel(), _stream())
    _capi.check(rc, "lgc_bpr_loss_grad")
    return loss3
