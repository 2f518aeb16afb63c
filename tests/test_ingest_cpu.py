"""Ingest plumbing against the reference's own pandas code: `relabelling`,
`interact_matrix` and `pos_item_list` (src/utils_v2.py:40-103), run here through the reference shim.
Index work: bit-exact."""
import numpy as np
import pandas as pd
import pytest
import torch

from gnn_ecommerce_b200 import ingest, synth
from oracle import reference_shim

pytestmark = pytest.mark.skipif(not reference_shim.reference_available(), reason="reference sources not available")


def _frame(seed=1):
    g = synth.make_graph(3000, 400, 30_000, seed=seed)
    rng = np.random.default_rng(seed)
    un = rng.permutation(5_000_000)[:g.n_users] + 17
    inn = rng.permutation(800_000)[:g.n_items] + 3
    return pd.DataFrame({"user_id": un[g.user], "item_id": inn[g.item - g.n_users], "weight": g.weight.astype(np.float64)})


def test_relabel_matches_label_encoder_and_transform_rejects_unseen_ids():
    _, ref_utils = reference_shim.load_reference()
    df = _frame()
    val = df.sample(500, random_state=0).copy()
    n_users, n_items, train, val2, _ = ref_utils.relabelling(df.copy(), val, None)
    ids = ingest.relabel(torch.from_numpy(df["user_id"].to_numpy()), torch.from_numpy(df["item_id"].to_numpy()))
    assert (ids.n_users, ids.n_items) == (n_users, n_items)
    assert np.array_equal(ids.user_idx.numpy(), train["user_id_idx"].to_numpy())
    assert np.array_equal(ids.item_idx.numpy(), train["item_id_idx"].to_numpy())
    assert np.array_equal(ids.transform_users(torch.from_numpy(val["user_id"].to_numpy())).numpy(),
                          val2["user_id_idx"].to_numpy())
    with pytest.raises(ValueError):
        ids.transform_items(torch.tensor([10**12]))


def test_seen_lists_equal_the_dense_interaction_mask_and_positive_lists_equal_pos_item_list():
    _, ref_utils = reference_shim.load_reference()
    df = _frame(2)
    n_users, n_items, train, _, _ = ref_utils.relabelling(df.copy())
    u = torch.from_numpy(train["user_id_idx"].to_numpy())
    i = torch.from_numpy(train["item_id_idx"].to_numpy())
    w = torch.from_numpy(train["weight"].to_numpy()).float()
    dense = ref_utils.interact_matrix(train, n_users, n_items)
    users = torch.tensor([0, 5, 77, 1234, n_users - 1])
    want = torch.index_select(dense, 0, users).to_dense()
    seen = ingest.seen_lists(u, i, w, n_users, n_items, users)
    got = torch.zeros_like(want)
    for r in range(users.numel()):
        got[r, seen.items[seen.ptr[r]:seen.ptr[r + 1]]] = 1.0
    assert torch.equal(got, (want > 0).float())
    full = ingest.seen_lists(u, i, w, n_users, n_items)
    assert int(full.ptr[-1]) == int((dense.coalesce().values() > 0).sum())
    pos = ref_utils.pos_item_list(train)
    pu, ptr, items = ingest.positive_lists(u, i, w, n_users)
    assert np.array_equal(pu.numpy(), pos["user_id_idx"].to_numpy())
    for r in (0, 3, len(pos) - 1):
        assert items[ptr[r]:ptr[r + 1]].tolist() == pos["item_id_idx_list"].iloc[r]
