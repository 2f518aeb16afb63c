"""ctypes binding of liblgc_b200.so (the C ABI declared in include/lgc_b200.h).

This is the whole host<->kernel boundary: raw device pointers, sizes and a CUDA stream handle.
There is NO fallback: if the shared library is missing or a call fails, a `RuntimeError` is
raised -- the CPU oracle under `oracle/` is test infrastructure and is never imported here.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_PKG = os.path.dirname(os.path.abspath(__file__))
# LGC_B200_LIB: another build of the same library (A/B experiments under profiles/experiments)
LIB_PATH = os.environ.get("LGC_B200_LIB") or os.path.join(_PKG, "liblgc_b200.so")

c_i64, c_i32, c_f32, c_f64, c_vp, c_sz = C.c_int64, C.c_int32, C.c_float, C.c_double, C.c_void_p, C.c_size_t


class GraphInfo(C.Structure):
    _fields_ = [("num_nodes", c_i64), ("nnz", c_i64), ("is_symmetric", c_i32),
                ("light_max_degree", c_i32), ("num_heavy_rows", c_i64), ("num_chunks", c_i64),
                ("num_split_rows", c_i64), ("rowptr", c_vp), ("src", c_vp), ("eid", c_vp),
                ("w_hat", c_vp), ("deg", c_vp), ("dis", c_vp)]


class PlanInfo(C.Structure):
    _fields_ = [("has_plan", c_i32), ("sweep_slots_per_unit", c_i32), ("sweep_units", c_i64),
                ("sweep_rows", c_i64), ("sweep_edges", c_i64), ("sweep_sources", c_i64),
                ("sweep_pieces_split_rows", c_i64), ("sweep_partial_slots", c_i64),
                ("sweep_iterations", c_i64), ("sweep_windows", c_i64),
                ("rows_rows", c_i64), ("rows_edges", c_i64), ("rows_sources", c_i64)]


class TrainStepArgs(C.Structure):
    _fields_ = [("ld", c_i32), ("num_layers", c_i32), ("h_alpha", C.POINTER(c_f32)),
                ("batch", c_i64), ("users", c_vp), ("pos", c_vp), ("neg", c_vp),
                ("decay", c_f64), ("lr", c_f64), ("beta1", c_f64), ("beta2", c_f64), ("eps", c_f64),
                ("step", c_i64), ("e0", c_vp), ("m", c_vp), ("v", c_vp), ("loss3", c_vp),
                ("workspace", c_vp), ("workspace_bytes", c_sz)]


class SpmmEpilogue(C.Structure):
    _fields_ = [("mode", c_i32), ("a0", c_f32), ("a1", c_f32), ("scale", c_f32), ("beta", c_f32),
                ("y", c_vp), ("acc", c_vp), ("xrow", c_vp), ("addend", c_vp), ("p", c_vp), ("m", c_vp),
                ("v", c_vp), ("lr", c_f64), ("beta1", c_f64), ("beta2", c_f64), ("eps", c_f64),
                ("step", c_i64), ("adam_scalars", c_vp), ("hist", c_vp * 6), ("ah", c_f32 * 6),
                ("n_hist", c_i32)]


class ScoreTopkArgs(C.Structure):
    _fields_ = [("d", c_i32), ("ld_user", c_i32), ("ld_item", c_i32), ("k", c_i32),
                ("n_users", c_i64), ("n_items", c_i64), ("user_emb", c_vp), ("item_emb", c_vp),
                ("user_ids", c_vp), ("seen_ptr", c_vp), ("seen_items", c_vp),
                ("topk_items", c_vp), ("topk_scores", c_vp), ("stats", c_vp),
                ("workspace", c_vp), ("workspace_bytes", c_sz)]


class PeerExchange(C.Structure):
    _fields_ = [("world", c_i32), ("rank", c_i32), ("bases", c_vp * 8), ("arena_bytes", c_sz), ("ctrl_off", c_sz),
                ("part", c_vp), ("n_rows", c_i64), ("ld", c_i32), ("timeout_ms", c_i32)]


# name -> (restype, argtypes); mirrors include/lgc_b200.h one to one
_SIGNATURES = {
    "lgc_abi_version": (C.c_int, []),
    "lgc_last_error": (C.c_char_p, []),
    "lgc_ld_supported": (C.c_int, [C.c_int]),
    "lgc_launch_count": (C.c_longlong, []),
    "lgc_profile_enable": (C.c_int, [C.c_int]),
    "lgc_profile_read": (C.c_int, [C.POINTER(c_f64), C.POINTER(C.c_longlong), C.c_int]),
    "lgc_graph_build": (C.c_int, [c_i64, c_i64, c_vp, c_vp, C.c_int, c_vp, C.POINTER(c_vp)]),
    "lgc_graph_build_pairs": (C.c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, C.c_int, c_vp, C.POINTER(c_vp)]),
    "lgc_graph_build_rect": (C.c_int, [c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, C.POINTER(c_vp)]),
    "lgc_graph_destroy": (C.c_int, [c_vp]),
    "lgc_graph_get_info": (C.c_int, [c_vp, C.POINTER(GraphInfo)]),
    "lgc_graph_plan_info": (C.c_int, [c_vp, C.c_int, C.POINTER(PlanInfo)]),
    "lgc_spmm_workspace_bytes": (c_sz, [c_vp, C.c_int]),
    "lgc_spmm": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lgc_spmm_ex": (C.c_int, [c_vp, C.c_int, c_vp, C.POINTER(SpmmEpilogue), c_vp, c_sz, c_vp]),
    "lgc_epilogue_apply": (C.c_int, [c_i64, C.c_int, c_vp, C.POINTER(SpmmEpilogue), c_vp]),
    "lgc_propagate_workspace_bytes": (c_sz, [c_vp, C.c_int, C.c_int]),
    "lgc_propagate": (C.c_int, [c_vp, C.c_int, C.c_int, C.POINTER(c_f32), c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lgc_pair_scores": (C.c_int, [C.c_int, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "lgc_bpr_workspace_bytes": (c_sz, [c_i64]),
    "lgc_bpr_loss_grad": (C.c_int, [c_i64, C.c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_f64, c_f32,
                                    c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lgc_scatter_add_rows": (C.c_int, [c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp]),
    "lgc_adam_step": (C.c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_f64, c_f64, c_f64, c_f64, c_i64, c_vp]),
    "lgc_adam_scalars": (C.c_int, [c_f64, c_f64, c_f64, c_f64, c_i64, C.POINTER(c_f32)]),
    "lgc_adam_step_dev": (C.c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "lgc_train_step_workspace_bytes": (c_sz, [c_vp, C.c_int, C.c_int, c_i64]),
    "lgc_train_workspace_init": (C.c_int, [c_vp, C.c_int, C.c_int, c_i64, c_vp, c_sz, c_vp]),
    "lgc_train_step": (C.c_int, [c_vp, C.POINTER(TrainStepArgs), c_vp]),
    "lgc_sample_triples_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "lgc_sample_triples": (C.c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, C.c_uint64,
                                     C.c_uint64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lgc_debug_light_phases": (C.c_int, [c_vp]),
    "lgc_mark_mapk": (C.c_int, [c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "lgc_score_topk_workspace_bytes": (c_sz, [c_i64, c_i64, C.c_int, C.c_int]),
    "lgc_score_topk": (C.c_int, [C.POINTER(ScoreTopkArgs), c_vp]),
    "lgc_peer_arena_alloc": (C.c_int, [c_sz, C.POINTER(c_vp), c_vp]),
    "lgc_peer_arena_open": (C.c_int, [c_vp, C.POINTER(c_vp)]),
    "lgc_peer_arena_close": (C.c_int, [c_vp]),
    "lgc_peer_arena_free": (C.c_int, [c_vp]),
    "lgc_item_exchange": (C.c_int, [C.POINTER(PeerExchange), C.POINTER(SpmmEpilogue), c_vp]),
    "lgc_peer_exchange_status": (C.c_int, [C.POINTER(PeerExchange), C.POINTER(c_i32), C.POINTER(c_i64)]),
}

_lib: Optional[C.CDLL] = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    """Load the shared library once. Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m gnn_ecommerce_b200.build` "
                "(there is no CPU or PyTorch fallback for the LightGCN hot path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so lacks a symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().lgc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {rc}: {msg}")
