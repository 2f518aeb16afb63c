"""Import the reference's own `LightGCN` / `utils_v2` UNMODIFIED (TEST INFRASTRUCTURE).

Only works where `/root/reference` exists (the build container). It is used by
`tests/golden/make_golden.py` to emit the committed golden vectors and by CPU tests that
cross-check `oracle/port.py`; nothing that runs on the GPU box may call it.

`src/lightgcn.py:8-10` imports three absent third-party modules. They are stubbed in
`sys.modules` with the minimum the file touches:
* `torch_sparse.SparseTensor`   -- only used in an `isinstance` (`src/lightgcn.py:116`);
* `torch_geometric.typing.{Adj, OptTensor}` -- annotations only;
* `torch_geometric.nn.conv.LGConv` -- the restated operator of `oracle/lgconv.py`.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Optional

import torch

REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "lightgcn.py"))


def _install_stubs() -> None:
    from oracle.lgconv import LGConv

    if "torch_sparse" not in sys.modules:
        ts = types.ModuleType("torch_sparse")

        class SparseTensor:  # never instantiated on this path
            pass

        ts.SparseTensor = SparseTensor
        sys.modules["torch_sparse"] = ts
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tg_nn = types.ModuleType("torch_geometric.nn")
        tg_conv = types.ModuleType("torch_geometric.nn.conv")
        tg_typing = types.ModuleType("torch_geometric.typing")
        tg_conv.LGConv = LGConv
        tg_typing.Adj = torch.Tensor
        tg_typing.OptTensor = Optional[torch.Tensor]
        tg.nn, tg_nn.conv, tg.typing = tg_nn, tg_conv, tg_typing
        sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tg_nn,
                            "torch_geometric.nn.conv": tg_conv,
                            "torch_geometric.typing": tg_typing})


def load_reference():
    """Return the reference modules `(lightgcn, utils_v2)`, imported from /root/reference/src."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    try:
        ref_lightgcn = importlib.import_module("lightgcn")
        ref_utils = importlib.import_module("utils_v2")
    finally:
        sys.path.remove(REFERENCE_SRC)
    if not getattr(ref_lightgcn, "__file__", "").startswith("/root/reference"):
        raise RuntimeError("a different `lightgcn` module shadows the reference's")
    return ref_lightgcn, ref_utils
