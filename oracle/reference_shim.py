"""Import the reference's own `LightGCN` / `utils_v2` / `train_lightgcn` UNMODIFIED (TEST INFRASTRUCTURE).

Sources: `/root/reference/src` in the build container, else the staged copy `oracle/_ref/src`
(`oracle/stage_reference.py`; git-ignored, travels to the GPU box). Used by
`tests/golden/make_golden.py` to emit the committed golden vectors, by the tests that cross-check
`oracle/port.py` and drive the reference's own training loop over the drop-in module, and by the
CPU legs of `bench.py` (`cpu_baseline.kind = "reference"`). Never by the product path.

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

_STAGED_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")
REFERENCE_SRC = "/root/reference/src" if os.path.isfile("/root/reference/src/lightgcn.py") else _STAGED_SRC


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
    """Return the reference modules `(lightgcn, utils_v2)`, imported from REFERENCE_SRC."""
    if not reference_available():
        raise RuntimeError("neither /root/reference nor the staged copy oracle/_ref is present")
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    try:
        ref_lightgcn = importlib.import_module("lightgcn")
        ref_utils = importlib.import_module("utils_v2")
    finally:
        sys.path.remove(REFERENCE_SRC)
    if not os.path.abspath(getattr(ref_lightgcn, "__file__", "")).startswith(os.path.abspath(REFERENCE_SRC)):
        raise RuntimeError("a different `lightgcn` module shadows the reference's")
    return ref_lightgcn, ref_utils


def load_reference_trainer(lightgcn_module=None):
    """The reference's `train_lightgcn` module (its `TrainLightGCN` class, src/train_lightgcn.py:8-162),
    imported unmodified. `lightgcn_module`: what its `from lightgcn import LightGCN` resolves to --
    None = the reference's own `lightgcn.py`; pass a module exposing the drop-in `LightGCN` to run the
    reference's training / evaluation loop over the B200 implementation (the drop-in proof)."""
    ref_lightgcn, ref_utils = load_reference()
    saved = {k: sys.modules.get(k) for k in ("lightgcn", "utils_v2", "train_lightgcn")}
    sys.modules["lightgcn"] = lightgcn_module if lightgcn_module is not None else ref_lightgcn
    sys.modules["utils_v2"] = ref_utils
    sys.modules.pop("train_lightgcn", None)
    sys.path.insert(0, REFERENCE_SRC)
    try:
        mod = importlib.import_module("train_lightgcn")
    finally:
        sys.path.remove(REFERENCE_SRC)
        sys.modules.pop("train_lightgcn", None)
        for k in ("lightgcn", "utils_v2"):
            if saved[k] is not None:
                sys.modules[k] = saved[k]
    return mod


def reference_train_step(ref_utils, model, optimizer, edge_index, edge_weight, users, pos, neg, decay):
    """The body of `mini_batch_loop` (src/train_lightgcn.py:129-151) for pre-sampled triples, with the
    reference's own model methods and `utils_v2` helpers. Returns the three `.item()` reads."""
    optimizer.zero_grad()
    labels = ref_utils.batch_pos_neg_edges(users, pos, neg)
    out = model(edge_index, labels, edge_weight)
    size = len(users)
    bpr = model.recommendation_loss(out[:size], out[size:], 0) * size
    reg = ref_utils.regularization_loss(model.embedding.weight, size, users, pos, neg, decay)
    loss = bpr + reg
    loss.backward()
    optimizer.step()
    return bpr.item(), reg.item(), loss.item()
