"""Checkpoint wire format of the reference (`save_model`, `src/utils_v2.py:212-230`; loaded by
`src/train_lightgcn.py:64-70`, `src/inference_lightgcn.py:32-35`, `torchserve/lightgcn_handler.py:44-47`):

    {'timestamp', 'epoch', 'model_state_dict': {alpha, embedding.weight}, 'optimizer_state_dict',
     'precision', 'recall', 'hyperparams': {'latent_dim', 'n_layers', ...}}

`model_state_dict` always carries the LOGICAL `[num_nodes, embedding_dim]` table (device storage
may be padded, e.g. d = 90 -> 96 floats per row) and `optimizer_state_dict` is torch.optim.Adam's
format whether the optimiser is torch's Adam or `FusedBPRTrainer`, so files written here load into
the reference and files written by the reference load here -- including the Adam moments, which
the reference saves but never reloads (true resume)."""
from __future__ import annotations

from datetime import datetime
from typing import Optional

import torch


def checkpoint_dict(model, optimizer, precision, recall, epoch=None, hyperparams=None) -> dict:
    sd = {k: v.detach().clone().contiguous().cpu() for k, v in model.state_dict().items()}
    return {"timestamp": datetime.now().strftime("%Y-%m-%d %H:%M:%S"), "epoch": epoch, "model_state_dict": sd,
            "optimizer_state_dict": optimizer.state_dict(), "precision": precision, "recall": recall,
            "hyperparams": hyperparams}


def save_model(path, model, optimizer, precision, recall, epoch=None, hyperparams=None) -> None:
    """Same signature and file layout as the reference's `save_model`."""
    torch.save(checkpoint_dict(model, optimizer, precision, recall, epoch, hyperparams), path)


def load_model(path_or_dict, device, trainer_lr: Optional[float] = None):
    """Rebuild `(model, trainer, checkpoint)` from a reference-format checkpoint: the model from
    `hyperparams['latent_dim' | 'n_layers']` like `src/train_lightgcn.py:66-69`, plus a
    `FusedBPRTrainer` carrying the saved Adam moments and step count."""
    from .lightgcn import LightGCN
    from .trainer import FusedBPRTrainer
    ck = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu",
                                                                        weights_only=False)
    w = ck["model_state_dict"]["embedding.weight"]
    hp = ck.get("hyperparams") or {}
    dim = int(hp.get("latent_dim", w.shape[1]))
    layers = int(hp.get("n_layers", ck["model_state_dict"]["alpha"].numel() - 1))
    model = LightGCN(w.shape[0], dim, layers)
    model.load_state_dict(ck["model_state_dict"])
    model = model.to(device)
    opt_sd = ck.get("optimizer_state_dict")
    lr = trainer_lr if trainer_lr is not None else (opt_sd["param_groups"][0]["lr"] if opt_sd else 0.005)
    trainer = FusedBPRTrainer(model, lr=lr)
    if opt_sd:
        trainer.load_state_dict(opt_sd)
    return model, trainer, ck
