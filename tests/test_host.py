"""CPU tests of the host side: C-ABI symbol coverage, module surface, host-only logic."""
import os
import re

import numpy as np
import pandas as pd
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "lgc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lgc_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from gnn_ecommerce_b200 import _capi, build
    build.build()
    lib = _capi.lib()
    declared = _header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lgc_b200.h but not exported"
    assert sorted(_capi.exported_symbols()) == declared       # ctypes table mirrors the header
    assert lib.lgc_abi_version() == 1                          # host-only call, no GPU needed
    for ld, ok in ((64, 1), (96, 1), (80, 1), (128, 1), (100, 0), (66, 0)):
        assert lib.lgc_ld_supported(ld) == ok


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "gnn_ecommerce_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_module_surface_matches_the_reference():
    from gnn_ecommerce_b200 import BPRLoss, LGConv, LightGCN
    m = LightGCN(50, 24, 4)
    assert list(m.state_dict().keys()) == ["alpha", "embedding.weight"]
    assert m.embedding.weight.shape == (50, 24) and m.alpha.shape == (5,)
    assert torch.allclose(m.alpha, torch.full((5,), 0.2))
    assert len(m.convs) == 4 and all(isinstance(c, LGConv) for c in m.convs)
    assert repr(m) == "LightGCN(50, 24, num_layers=4)"
    for name in ("get_embedding", "forward", "predict_link", "recommend", "recommendK", "MARK_MAPK",
                 "link_pred_loss", "recommendation_loss", "reset_parameters"):
        assert callable(getattr(m, name))
    bound = np.sqrt(6.0 / (50 + 24))
    assert float(m.embedding.weight.abs().max()) <= bound
    m2 = LightGCN(50, 24, 2, alpha=torch.tensor([0.5, 0.3, 0.2]))
    assert torch.equal(m2.alpha, torch.tensor([0.5, 0.3, 0.2]))
    # BPRLoss scaling identical to the reference's (src/lightgcn.py:279-286)
    pos, neg = torch.tensor([1.0, 0.5, -0.2]), torch.tensor([0.3, 0.7, 0.1])
    want = -torch.nn.functional.logsigmoid(pos - neg).mean() / 3
    assert torch.allclose(BPRLoss(0)(pos, neg), want)
    assert torch.allclose(m.recommendation_loss(pos, neg, 0) * 3, want * 3)


def test_cpu_tensors_fail_loudly_instead_of_falling_back():
    from gnn_ecommerce_b200 import FusedBPRTrainer, LightGCN
    m = LightGCN(10, 16, 2)
    ei = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(RuntimeError, match="no CPU"):
        m.get_embedding(ei, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedBPRTrainer(m).step(ei, None, torch.tensor([0]), torch.tensor([1]), torch.tensor([1]), 1e-4)


def test_missing_library_fails_loudly(monkeypatch):
    from gnn_ecommerce_b200 import _capi
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/liblgc_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _capi.lib()


def test_padded_dim_and_seen_lists():
    from gnn_ecommerce_b200.graph import padded_dim
    from gnn_ecommerce_b200.scoring import as_seen_lists
    assert [padded_dim(d) for d in (16, 64, 80, 90, 96, 100)] == [16, 64, 80, 96, 96, 128]
    mask = torch.zeros(4, 7)
    mask[0, 3] = 1; mask[2, 0] = 1; mask[2, 6] = 1
    s = as_seen_lists(mask, 4, 7, "cpu")
    assert s.ptr.tolist() == [0, 1, 1, 3, 3] and s.items.tolist() == [3, 0, 6]
    s2 = as_seen_lists(mask.to_sparse(), 4, 7, "cpu")
    assert s2.ptr.tolist() == s.ptr.tolist() and s2.items.tolist() == s.items.tolist()
    with pytest.raises(NotImplementedError):
        as_seen_lists(mask * 2, 4, 7, "cpu")


def test_mark_mapk_semantics_match_the_oracle():
    from gnn_ecommerce_b200 import LightGCN
    from oracle import port
    m = LightGCN(10, 8, 1)
    pos = pd.DataFrame({"user_id_idx": [0, 3, 5], "item_id_idx_list": [[1, 2], [4], [7, 8, 9, 9]]})
    top = pd.DataFrame({"user_ID": [5, 0, 3], "top_rlvnt_itm": [[9, 1, 0], [2, 1, 5], [0, 1, 2]]})
    prec, rec, frame = m.MARK_MAPK(pos, top, 3)
    want = port.mark_mapk([[1, 2], [4], [7, 8, 9, 9]], np.array([[2, 1, 5], [0, 1, 2], [9, 1, 0]]), 3)
    assert (prec, rec) == pytest.approx(want)
    assert list(frame["recall"]) == [1.0, 0.0, 0.25]


def test_synth_sampler_semantics():
    from gnn_ecommerce_b200 import synth
    g = synth.make_graph(500, 80, 4000, seed=1)
    held = synth.make_heldout(g, 50)
    pl = synth.purchase_lists(g, held)
    u, p, n = synth.sample_triples(pl, 64, g.n_users, g.n_items, np.random.default_rng(0))
    assert len(set(u.tolist())) == 64                                     # distinct users
    purchases = set(zip(g.user[g.weight == 1].tolist(), g.item[g.weight == 1].tolist()))
    assert all((a, b) in purchases for a, b in zip(u, p))                # positives are purchases
    assert all((a, b) not in purchases for a, b in zip(u, n))            # negatives are not
    assert n.min() >= g.n_users and n.max() < g.num_nodes
    ptr, items = synth.seen_lists(g, held.users)
    for i, usr in enumerate(held.users[:10]):
        want = sorted(b - g.n_users for a, b in purchases if a == usr)
        assert items[ptr[i]:ptr[i + 1]].tolist() == want
