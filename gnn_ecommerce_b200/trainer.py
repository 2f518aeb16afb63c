"""Fused BPR training step: the body of `TrainLightGCN.mini_batch_loop`
(reference `src/train_lightgcn.py:129-151`) as ONE C-ABI call per mini-batch.

    trainer = FusedBPRTrainer(model, lr=0.005)              # replaces torch.optim.Adam(...)
    bpr, reg, total = trainer.step(edge_index, edge_weight, users, pos, neg, decay)

Forward (K fused SpMM layers), BPR + L2 loss, backward (K SpMM layers, Horner form) and the dense
Adam update (fused into the last backward layer) run back to back on the current stream with no
host synchronisation; the three losses come back as a device tensor (read them once per epoch,
not three `.item()` syncs per step like the reference).

`state_dict()` / `load_state_dict()` speak torch.optim.Adam's format so the reference's
`save_model` (`src/utils_v2.py:214-232`) keeps working.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _capi, ops
from .graph import _ptr, _stream, padded_dim


class FusedBPRTrainer:
    def __init__(self, model, lr: float = 0.005, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8):
        self.model = model
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0
        self.m: Optional[Tensor] = None
        self.v: Optional[Tensor] = None
        self._ws = None
        self._ws_key = None
        self._lib = _capi.lib()

    # ------------------------------------------------------------------ state
    def _ensure_state(self) -> Tensor:
        """Make `model.embedding.weight` a [N, d] view of a padded [N, ld] table (zero-copy when
        d is itself a supported width) and allocate the Adam moments alongside."""
        w = self.model.embedding.weight
        if not w.is_cuda:
            raise RuntimeError("FusedBPRTrainer: move the model to a CUDA device first "
                               "(no CPU fallback)")
        d, ld = w.size(1), padded_dim(w.size(1))
        t = w.data
        ok = (t.stride(1) == 1 and t.stride(0) == ld and t.data_ptr() % 16 == 0
              and (ld == d or ops._pad_is_zero_view(t, ld)))
        if not ok:
            padded = ops.new_table(w.size(0), d, w.device)
            padded.copy_(t)
            w.data = padded                      # same Parameter object, padded storage
        table = w.data
        rows = ops.full_rows(table)
        if self.m is None or self.m.shape != rows.shape or self.m.device != rows.device:
            m_old, v_old = self.m, self.v
            self.m, self.v = torch.zeros_like(rows), torch.zeros_like(rows)
            if m_old is not None and m_old.size(0) == rows.size(0):
                dd = min(m_old.size(1), rows.size(1))
                self.m[:, :dd] = m_old[:, :dd].to(rows.device)
                self.v[:, :dd] = v_old[:, :dd].to(rows.device)
        return rows

    def _workspace(self, g, ld: int, k: int, batch: int) -> Tensor:
        key = (id(g), ld, k, batch)
        if self._ws_key != key:
            nbytes = self._lib.lgc_train_step_workspace_bytes(g.handle, ld, k, batch)
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=g.device)
            rc = self._lib.lgc_train_workspace_init(g.handle, ld, k, batch, _ptr(self._ws),
                                                    self._ws.numel(), _stream())
            _capi.check(rc, "lgc_train_workspace_init")
            self._ws_key = key
        return self._ws

    # ------------------------------------------------------------------ one mini-batch
    def step(self, edge_index: Tensor, edge_weight: Optional[Tensor], users: Tensor, pos: Tensor,
             neg: Tensor, decay: float) -> Tensor:
        """Returns a device tensor [bpr_loss, reg_loss, total] (fp32)."""
        model = self.model
        rows = self._ensure_state()
        g = model.graph(edge_index, edge_weight)
        if not g.is_symmetric:
            raise RuntimeError("the fused step needs a symmetric graph (both edge directions with "
                               "equal weights, as `df_to_graph` builds it); use autograd instead")
        ld, k = rows.stride(0), model.num_layers
        batch = users.numel()
        users, pos, neg = (t.to(device=rows.device, dtype=torch.int64).contiguous()
                           for t in (users, pos, neg))
        ws = self._workspace(g, ld, k, batch)
        alpha = model.alpha_host()
        loss3 = torch.empty(3, dtype=torch.float32, device=rows.device)
        args = _capi.TrainStepArgs(
            ld=ld, num_layers=k, h_alpha=(C.c_float * len(alpha))(*alpha), batch=batch,
            users=_ptr(users), pos=_ptr(pos), neg=_ptr(neg), decay=float(decay), lr=self.lr,
            beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, step=self.step_count + 1,
            e0=_ptr(rows), m=_ptr(self.m), v=_ptr(self.v), loss3=_ptr(loss3),
            workspace=_ptr(ws), workspace_bytes=ws.numel())
        with torch.cuda.device(rows.device):
            rc = self._lib.lgc_train_step(g.handle, C.byref(args), _stream())
        if rc != 0:
            self._ws_key = None          # a failed step may leave gradient rows behind: re-zero the workspace next time
        _capi.check(rc, "lgc_train_step")
        self.step_count += 1             # only a completed step advances the Adam bias correction
        model._weights_epoch = getattr(model, "_weights_epoch", 0) + 1     # invalidates cached_embedding
        return loss3

    def zero_grad(self, set_to_none: bool = True) -> None:   # API symmetry with torch.optim
        pass

    # ------------------------------------------------------------------ torch.optim.Adam format
    def state_dict(self):
        d = self.model.embedding_dim
        state = {}
        if self.m is not None:
            state[0] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.m[:, :d].contiguous(),
                        "exp_avg_sq": self.v[:, :d].contiguous()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                 "differentiable": False, "fused": None, "params": [0]}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd) -> None:
        group = sd["param_groups"][0]
        self.lr, self.eps = float(group["lr"]), float(group["eps"])
        self.betas = (float(group["betas"][0]), float(group["betas"][1]))
        st = sd["state"].get(0)
        if st is not None:
            rows = self._ensure_state()
            d = self.model.embedding_dim
            self.step_count = int(float(st["step"]))
            self.m.zero_(); self.v.zero_()
            self.m[:, :d] = st["exp_avg"].to(rows.device)
            self.v[:, :d] = st["exp_avg_sq"].to(rows.device)
