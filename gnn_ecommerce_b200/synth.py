"""Seeded synthetic Cosmetics-Shop-shaped inputs for the LightGCN hot path.

The real dataset of the reference is a DVC pointer to an unreachable remote
(reference `.dvc/config:4`), so every test and benchmark runs on synthetic
power-law bipartite graphs whose shape follows SURVEY.md section 8(d):

* users ``0..n_users-1``; items ``n_users..N-1`` (offset convention of reference
  `src/utils_v2.py:128`);
* user degree ~ truncated power law (min 1, long tail), item popularity ~ Zipf with a few
  hub items; every user and every item owns at least one train edge (the reference derives
  its node set from the train frame, `src/utils_v2.py:48-60`);
* weights drawn from the reference's value set (`config.yaml:10`,
  `notebooks/1.data_preprocessing.ipynb`): 1.0 marks a purchase (positive / seen item);
* directed edge list in the layout of `df_to_graph` (`src/utils_v2.py:146-165`): first the E
  user->item entries in frame order, then the E item->user entries in the same order.

Pure numpy, no torch: usable from tests, bench.py and the oracle alike.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np

WEIGHT_VALUES = np.array([0.01, 0.02, 0.1, 0.11, 0.5, 1.0], dtype=np.float32)
WEIGHT_PROBS = np.array([0.60, 0.10, 0.10, 0.05, 0.05, 0.10])

# name -> (n_users, n_items, E undirected train edges, d, K)
CONFIGS: Dict[str, Tuple[int, int, int, int, int]] = {
    "c1": (10_000, 2_000, 100_000, 64, 3),
    "c2": (1_600_000, 54_000, 5_000_000, 64, 3),
    "c3": (1_600_000, 54_000, 5_000_000, 90, 5),
    "c5": (16_000_000, 500_000, 200_000_000, 64, 3),
}


@dataclass
class BipartiteGraph:
    """Train frame of the reference after `prepare_val_test`, as plain arrays."""
    n_users: int
    n_items: int
    user: np.ndarray      # int64 [E], user_id_idx
    item: np.ndarray      # int64 [E], item_id_idx already offset by n_users
    weight: np.ndarray    # float32 [E]

    @property
    def num_nodes(self) -> int:
        return self.n_users + self.n_items

    @property
    def num_edges(self) -> int:
        return int(self.user.shape[0])

    def edge_index(self) -> np.ndarray:
        """int64 [2, 2E] in the `df_to_graph` layout (reference `src/utils_v2.py:153-158`)."""
        return np.stack((np.concatenate([self.user, self.item]),
                         np.concatenate([self.item, self.user])))

    def edge_weight(self) -> np.ndarray:
        """float32 [2E] (reference `src/utils_v2.py:160-163`)."""
        return np.concatenate([self.weight, self.weight])


def _user_degrees(rng: np.random.Generator, n_users: int, n_edges: int, cap: int) -> np.ndarray:
    mean_extra = n_edges / n_users - 1.0
    if mean_extra < 0:
        raise ValueError("need at least one edge per user")
    shape = 1.6                                   # Lomax tail exponent: median << mean
    raw = rng.pareto(shape, n_users) * (mean_extra * (shape - 1.0))
    deg = 1 + np.minimum(np.floor(raw).astype(np.int64), cap - 1)
    diff = n_edges - int(deg.sum())
    while diff != 0:                              # nudge the total to exactly n_edges
        if diff > 0:
            idx = rng.integers(0, n_users, size=diff)
            np.add.at(deg, idx, 1)
        else:
            cand = np.flatnonzero(deg > 1)
            take = rng.choice(cand, size=min(-diff, cand.size), replace=False)
            deg[take] -= 1
        np.minimum(deg, cap, out=deg)
        diff = n_edges - int(deg.sum())
    return deg


def _zipf_cdf(n_items: int, exponent: float = 1.05, shift: float = 8.0) -> np.ndarray:
    p = 1.0 / np.power(np.arange(n_items, dtype=np.float64) + shift, exponent)
    cdf = np.cumsum(p)
    return cdf / cdf[-1]


def make_graph(n_users: int, n_items: int, n_edges: int, seed: int = 42,
               degree_cap: int = 20_000) -> BipartiteGraph:
    """Seeded power-law bipartite train graph with E unique (user, item) pairs."""
    if n_edges < max(n_users, n_items):
        raise ValueError("n_edges must cover every user and every item once")
    rng = np.random.default_rng(seed)
    degree_cap = max(1, min(degree_cap, n_items // 4))
    deg = _user_degrees(rng, n_users, n_edges, degree_cap)
    slot_user = np.repeat(np.arange(n_users, dtype=np.int64), deg)        # [E]
    cdf = _zipf_cdf(n_items)
    item_of_rank = rng.permutation(n_items)        # popularity rank -> item id

    # One forced edge per item (random distinct slots) so that no item is isolated.
    forced_slots = rng.choice(n_edges, size=n_items, replace=False)
    slot_item = np.full(n_edges, -1, dtype=np.int64)
    slot_item[forced_slots] = rng.permutation(n_items)
    forced = np.zeros(n_edges, dtype=bool)
    forced[forced_slots] = True

    pending = np.flatnonzero(slot_item < 0)
    rounds = 0
    while pending.size:
        if rounds < 6:
            draw = item_of_rank[np.searchsorted(cdf, rng.random(pending.size))]
        else:                                      # stubborn duplicates: uniform redraw
            draw = rng.integers(0, n_items, size=pending.size)
        slot_item[pending] = draw
        key = slot_user * n_items + slot_item
        # keep the first occurrence of every (user, item); forced slots sort first
        order = np.lexsort((~forced, key))
        sk = key[order]
        dup = np.empty(n_edges, dtype=bool)
        dup[0] = False
        dup[1:] = sk[1:] == sk[:-1]
        pending = order[dup]
        rounds += 1

    perm = rng.permutation(n_edges)                # frame row order
    user = slot_user[perm]
    item = slot_item[perm] + n_users
    weight = WEIGHT_VALUES[rng.choice(WEIGHT_VALUES.size, size=n_edges, p=WEIGHT_PROBS)]
    return BipartiteGraph(n_users, n_items, user, item, weight.astype(np.float32))


def make_config_graph(name: str, seed: int = 42) -> BipartiteGraph:
    n_users, n_items, n_edges, _, _ = CONFIGS[name]
    return make_graph(n_users, n_items, n_edges, seed)


@dataclass
class PurchaseLists:
    """CSR view of the reference's `train_pos_list_df` (`src/utils_v2.py:64-89`)."""
    users: np.ndarray        # int64 [P] purchasers (users with >=1 weight==1 train edge), sorted
    pos_ptr: np.ndarray      # int64 [P+1]
    pos_items: np.ndarray    # int64 offset item ids, frame order inside a user
    ign_ptr: np.ndarray      # int64 [P+1]
    ign_items: np.ndarray    # int64 sorted offset item ids (train U val U test positives)


def purchase_lists(g: BipartiteGraph, heldout: "HeldOut | None" = None) -> PurchaseLists:
    mask = g.weight == np.float32(1.0)
    u, it = g.user[mask], g.item[mask]
    order = np.argsort(u, kind="stable")
    u, it = u[order], it[order]
    users, counts = np.unique(u, return_counts=True)
    pos_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    iu, ii = u, it
    if heldout is not None and heldout.users.size:
        hu = np.repeat(heldout.users, np.diff(heldout.ptr))
        keep = np.isin(hu, users)
        iu = np.concatenate([u, hu[keep]])
        ii = np.concatenate([it, heldout.items[keep] + g.n_users])
    key = np.unique(iu * (g.n_users + g.n_items) + ii)
    iu, ii = key // (g.n_users + g.n_items), key % (g.n_users + g.n_items)
    ign_counts = np.bincount(np.searchsorted(users, iu), minlength=users.size)
    ign_ptr = np.concatenate([[0], np.cumsum(ign_counts)]).astype(np.int64)
    return PurchaseLists(users.astype(np.int64), pos_ptr, it.astype(np.int64), ign_ptr,
                         ii.astype(np.int64))


def sample_triples(p: PurchaseLists, batch_size: int, n_users: int, n_items: int,
                   rng: np.random.Generator) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Seeded restatement of the semantics of `batch_loader` (reference `src/utils_v2.py:168-181`):
    distinct users among purchasers, positive uniform over the user's train purchases, negative
    uniform over items outside the user's ignore list (rejection sampling)."""
    if batch_size > p.users.size:
        raise ValueError("Sample larger than population")   # random.sample's own error
    sel = rng.choice(p.users.size, size=batch_size, replace=False)
    users = p.users[sel]
    lo, hi = p.pos_ptr[sel], p.pos_ptr[sel + 1]
    pos = p.pos_items[lo + (rng.random(batch_size) * (hi - lo)).astype(np.int64)]
    neg = np.empty(batch_size, dtype=np.int64)
    todo = np.arange(batch_size)
    while todo.size:
        cand = rng.integers(0, n_items, size=todo.size) + n_users
        bad = np.zeros(todo.size, dtype=bool)
        for j, t in enumerate(todo):               # ignore lists are tiny; host-side sampler
            s = sel[t]
            seg = p.ign_items[p.ign_ptr[s]:p.ign_ptr[s + 1]]
            k = np.searchsorted(seg, cand[j])
            bad[j] = k < seg.size and seg[k] == cand[j]
        neg[todo[~bad]] = cand[~bad]
        todo = todo[bad]
    return users.astype(np.int64), pos.astype(np.int64), neg.astype(np.int64)


@dataclass
class HeldOut:
    """Evaluation users and their held-out purchases (`val_pos_list_df` of the reference)."""
    users: np.ndarray        # int64 [U_eval], sorted, distinct
    ptr: np.ndarray          # int64 [U_eval+1]
    items: np.ndarray        # int64, un-offset item ids 0..n_items-1


def make_heldout(g: BipartiteGraph, n_eval_users: int, seed: int = 45,
                 mean_items: float = 1.5) -> HeldOut:
    """Held-out purchases for recall@k: pairs absent from the train graph, drawn with the same
    item popularity law. Mirrors what `sync_nodes` + `pos_item_list` leave of a 2.5 % split
    (reference `src/utils_v2.py:20-37,64-73`)."""
    rng = np.random.default_rng(seed)
    users = np.sort(rng.choice(g.n_users, size=min(n_eval_users, g.n_users), replace=False))
    cnt = 1 + rng.poisson(mean_items - 1.0, size=users.size)
    cdf = _zipf_cdf(g.n_items)
    item_of_rank = np.random.default_rng(seed + 1).permutation(g.n_items)
    hu = np.repeat(users, cnt)
    hi = item_of_rank[np.searchsorted(cdf, rng.random(hu.size))]
    train_key = g.user * g.n_items + (g.item - g.n_users)
    key = np.unique(hu * g.n_items + hi)
    key = key[~np.isin(key, train_key)]
    hu, hi = key // g.n_items, key % g.n_items
    users, cnt = np.unique(hu, return_counts=True)
    ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    return HeldOut(users.astype(np.int64), ptr, hi.astype(np.int64))


def seen_lists(g: BipartiteGraph, users: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """CSR (ptr int64 [U+1], sorted distinct un-offset item ids) of the train purchases of
    `users` -- the sparse form of the dense mask the reference builds with `interact_matrix`
    + `index_select(...).to_dense()` (`src/utils_v2.py:92-103,137-138`)."""
    mask = g.weight == np.float32(1.0)
    key = np.unique(g.user[mask] * g.n_items + (g.item[mask] - g.n_users))
    ku, ki = key // g.n_items, key % g.n_items
    users = np.asarray(users, dtype=np.int64)
    lo = np.searchsorted(ku, users, side="left")
    hi = np.searchsorted(ku, users, side="right")
    cnt = hi - lo
    ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    idx = np.repeat(lo - ptr[:-1], cnt) + np.arange(ptr[-1])
    return ptr, ki[idx].astype(np.int64)
