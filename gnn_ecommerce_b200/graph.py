"""Device graph handle: COO `edge_index` / `edge_weight` (the `df_to_graph` layout of reference
`src/utils_v2.py:146-165`) -> destination-major CSR + `gcn_norm` weights, built ONCE per graph.

The reference passes the same device tensors to `model(...)` on every step
(`src/train_lightgcn.py:35-37,138`) and PyG re-derives the normalisation inside every `LGConv`
call; here the derived CSR is cached per (storage pointer, shape, version) of the two tensors.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _capi

SUPPORTED_LD = (4, 8, 16, 32, 48, 64, 80, 96, 128, 160, 192, 256)


def padded_dim(d: int) -> int:
    """Smallest supported row width (floats) >= d; rows are runs of 128-bit loads."""
    for ld in SUPPORTED_LD:
        if ld >= d:
            return ld
    raise ValueError(f"embedding_dim {d} > {SUPPORTED_LD[-1]} is not supported")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Graph:
    """Owns an `lgc_graph_t*`."""

    def __init__(self, edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int,
                 normalize: bool = True):
        if not edge_index.is_cuda:
            raise RuntimeError("gnn_ecommerce_b200 runs on CUDA tensors only (no CPU fallback)")
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must be an int64 tensor of shape [2, nnz]")
        ei = edge_index.contiguous()
        ew = None
        if edge_weight is not None:
            if edge_weight.numel() != ei.size(1):
                raise ValueError("edge_weight must have one entry per edge")
            ew = edge_weight.to(torch.float32).contiguous()
        self._lib = _capi.lib()
        self.device = ei.device
        self.num_nodes = int(num_nodes)
        self.nnz = int(ei.size(1))
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self._lib.lgc_graph_build(self.num_nodes, self.nnz, _ptr(ei), _ptr(ew),
                                           1 if normalize else 0, _stream(), C.byref(handle))
        _capi.check(rc, "lgc_graph_build")
        self._adopt(handle)
        self._edge_index, self._edge_weight = ei, ew          # kept for the transpose build

    def _adopt(self, handle) -> None:
        self.handle = handle
        info = _capi.GraphInfo()
        _capi.check(self._lib.lgc_graph_get_info(self.handle, C.byref(info)), "lgc_graph_get_info")
        self.info = info
        self.is_symmetric = bool(info.is_symmetric)
        self._edge_index, self._edge_weight = None, None
        self._transpose: Optional["Graph"] = None
        self._ws = {}

    @classmethod
    def from_interactions(cls, user_idx: Tensor, item_idx: Tensor, weight: Optional[Tensor],
                          num_nodes: int, normalize: bool = True) -> "Graph":
        """The graph of `df_to_graph(train_df, True)` (reference `src/utils_v2.py:146-165`) built
        from the frame's two id columns directly: `user_idx` and `item_idx` (already offset by
        n_users, `src/utils_v2.py:128`) are int64 device tensors of E entries. Bit-identical to
        `Graph(edge_index, edge_weight, ...)` on df_to_graph's output, without materialising the
        `[2, 2E]` int64 COO. Pass the returned object wherever the module takes `edge_index`."""
        if not user_idx.is_cuda:
            raise RuntimeError("gnn_ecommerce_b200 runs on CUDA tensors only (no CPU fallback)")
        if user_idx.dtype != torch.int64 or item_idx.dtype != torch.int64 or user_idx.shape != item_idx.shape \
                or user_idx.dim() != 1:
            raise ValueError("user_idx / item_idx must be int64 vectors of equal length")
        a, b = user_idx.contiguous(), item_idx.to(user_idx.device).contiguous()
        ew = None
        if weight is not None:
            if weight.numel() != a.numel():
                raise ValueError("weight must have one entry per interaction")
            ew = weight.to(device=a.device, dtype=torch.float32).contiguous()
        self = cls.__new__(cls)
        self._lib = _capi.lib()
        self.device = a.device
        self.num_nodes = int(num_nodes)
        self.nnz = 2 * int(a.numel())
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self._lib.lgc_graph_build_pairs(self.num_nodes, int(a.numel()), _ptr(a), _ptr(b), _ptr(ew),
                                                 1 if normalize else 0, _stream(), C.byref(handle))
        _capi.check(rc, "lgc_graph_build_pairs")
        self._adopt(handle)
        return self

    # ---- enough of the tensor surface for a Graph to stand in for `edge_index` in the module API
    is_cuda = True
    _version = 0

    def data_ptr(self) -> int:
        return int(self.handle.value or 0)

    @property
    def shape(self):
        return (2, self.nnz)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                self._lib.lgc_graph_destroy(h)
            except Exception:  # interpreter shutdown
                pass

    # ---- views of the handle's device arrays (tests / diagnostics) -----------------------
    def _view(self, ptr: int, n: int, dtype: torch.dtype) -> Tensor:
        return _from_device_ptr(ptr, n, dtype, self.device)

    def arrays(self):
        i = self.info
        return {"rowptr": self._view(i.rowptr, self.num_nodes + 1, torch.int32),
                "src": self._view(i.src, self.nnz, torch.int32),
                "eid": self._view(i.eid, self.nnz, torch.int32),
                "w_hat": self._view(i.w_hat, self.nnz, torch.float32),
                "deg": self._view(i.deg, self.num_nodes, torch.float32),
                "dis": self._view(i.dis, self.num_nodes, torch.float32)}

    def w_hat_edge_order(self) -> Tensor:
        a = self.arrays()
        w = torch.empty(self.nnz, dtype=torch.float32, device=self.device)
        w[a["eid"].long()] = a["w_hat"]
        return w

    def transpose(self) -> "Graph":
        """A_hat^T as its own handle (only needed for the backward of a NON-symmetric graph;
        the reference's graphs are symmetric by construction, `src/utils_v2.py:155-158`)."""
        if self.is_symmetric:
            return self
        if self._transpose is None:
            if self._edge_index is None:
                raise RuntimeError("a graph built from interaction pairs is symmetric by construction")
            self._transpose = Graph(self._edge_index.flip(0), self.w_hat_edge_order(),
                                    self.num_nodes, normalize=False)
        return self._transpose

    def workspace(self, key, nbytes: int) -> Tensor:
        """Per-graph scratch (uint8), grown on demand and reused across calls."""
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf


class _DeviceArray:
    """Minimal `__cuda_array_interface__` carrier for a raw device pointer."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2}


_TYPESTR = {torch.int32: "<i4", torch.float32: "<f4", torch.int64: "<i8"}


def _from_device_ptr(ptr: int, n: int, dtype: torch.dtype, device: torch.device) -> Tensor:
    """Copy `n` elements starting at a raw device pointer into a fresh torch tensor."""
    if n == 0:
        return torch.empty(0, dtype=dtype, device=device)
    alias = torch.as_tensor(_DeviceArray(ptr, n, _TYPESTR[dtype]), device=device)
    return alias.clone()


_CACHE: "OrderedDict[Tuple, Graph]" = OrderedDict()
_CACHE_SIZE = 4


def graph_for(edge_index, edge_weight: Optional[Tensor], num_nodes: int,
              normalize: bool = True) -> Graph:
    if isinstance(edge_index, Graph):                      # a prebuilt graph passed where edge_index goes
        if edge_index.num_nodes != int(num_nodes):
            raise ValueError("the graph was built for a different number of nodes")
        return edge_index
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version,
           None if edge_weight is None else (edge_weight.data_ptr(), edge_weight._version,
                                             edge_weight.dtype),
           int(num_nodes), bool(normalize), edge_index.device.index)
    g = _CACHE.get(key)
    if g is None:
        g = Graph(edge_index, edge_weight, num_nodes, normalize)
        _CACHE[key] = g
        while len(_CACHE) > _CACHE_SIZE:
            _CACHE.popitem(last=False)
    else:
        _CACHE.move_to_end(key)
    return g


def clear_graph_cache() -> None:
    _CACHE.clear()
