"""gnn_ecommerce_b200 -- B200-native LightGCN hot path (propagation, BPR step, top-k scoring)
behind the `LightGCN` module API of happykygo/GNN-eCommerce (`src/lightgcn.py`).

Python/PyTorch is plumbing only (device memory, streams, autograd glue); the arithmetic is in
hand-written sm_100a CUDA behind the C ABI of `include/lgc_b200.h` (`liblgc_b200.so`).
Importing the package does not load CUDA; the first kernel call does, and fails loudly when the
library has not been built.
"""
from . import synth  # noqa: F401  (numpy only)

__all__ = ["LightGCN", "LGConv", "BPRLoss", "FusedBPRTrainer", "SeenLists", "score_topk", "synth",
           "DeviceSampler", "save_model", "load_model", "make_sharded_trainer", "LightGCNService"]


def __getattr__(name):
    if name in ("LightGCN", "LGConv", "BPRLoss"):
        from . import lightgcn
        return getattr(lightgcn, name)
    if name == "FusedBPRTrainer":
        from .trainer import FusedBPRTrainer
        return FusedBPRTrainer
    if name in ("SeenLists", "score_topk"):
        from . import scoring
        return getattr(scoring, name)
    if name == "DeviceSampler":
        from .sampler import DeviceSampler
        return DeviceSampler
    if name in ("save_model", "load_model"):
        from . import checkpoint
        return getattr(checkpoint, name)
    if name == "LightGCNService":
        from .serving import LightGCNService
        return LightGCNService
    if name == "make_sharded_trainer":
        from .sharded import make_sharded_trainer
        return make_sharded_trainer
    raise AttributeError(name)
