"""Device-side BPR sampler vs the semantics of the reference's `batch_loader`
(`src/utils_v2.py:168-181`). The reference sampler is unseeded, so parity is semantic and
distributional: every triple must be one the reference could have drawn, and the draw frequencies
must be uniform where the reference's are."""
import numpy as np
import pytest
import torch

from gnn_ecommerce_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sampler(g, seed=0):
    from gnn_ecommerce_b200.sampler import DeviceSampler
    pl = synth.purchase_lists(g)
    return pl, DeviceSampler.from_lists(pl.users, pl.pos_ptr, pl.pos_items, pl.ign_ptr, pl.ign_items, g.n_users,
                                        g.n_items, DEV, seed)


def test_triples_are_valid_and_reproducible():
    g = synth.make_graph(5000, 700, 60_000, seed=3)
    pl, s = _sampler(g, seed=11)
    pos_set = {int(u): set(pl.pos_items[pl.pos_ptr[i]:pl.pos_ptr[i + 1]].tolist()) for i, u in enumerate(pl.users)}
    ign_set = {int(u): set(pl.ign_items[pl.ign_ptr[i]:pl.ign_ptr[i + 1]].tolist()) for i, u in enumerate(pl.users)}
    batches = []
    for _ in range(5):
        u, p, n = (x.cpu().numpy() for x in s.sample(1024))
        batches.append((u, p, n))
        assert len(set(u.tolist())) == 1024                      # random.sample: distinct users
        for a, b, c in zip(u.tolist(), p.tolist(), n.tolist()):
            assert b in pos_set[a]                                # random.choice(user's purchases)
            assert g.n_users <= c < g.n_users + g.n_items and c not in ign_set[a]
    assert not np.array_equal(batches[0][0], batches[1][0])       # steps differ
    _, s2 = _sampler(g, seed=11)                                   # same seed, same sequence
    u, p, n = (x.cpu().numpy() for x in s2.sample(1024))
    assert np.array_equal(u, batches[0][0]) and np.array_equal(p, batches[0][1]) and np.array_equal(n, batches[0][2])


def test_sample_larger_than_population_raises_like_random_sample():
    g = synth.make_graph(300, 50, 2000, seed=1)
    pl, s = _sampler(g)
    with pytest.raises(ValueError, match="larger than population"):
        s.sample(len(pl.users) + 1)
    u, _, _ = s.sample(len(pl.users))                              # the whole population is allowed
    assert sorted(u.cpu().tolist()) == sorted(pl.users.tolist())


def test_draw_frequencies_are_uniform():
    g = synth.make_graph(4000, 200, 30_000, seed=5)
    pl, s = _sampler(g, seed=2)
    n_p = len(pl.users)
    cnt_u = np.zeros(g.n_users, np.int64)
    heavy = int(np.argmax(np.diff(pl.pos_ptr)))                    # the user with the most purchases
    hu = int(pl.users[heavy])
    h_pos = pl.pos_items[pl.pos_ptr[heavy]:pl.pos_ptr[heavy + 1]]
    h_ign = set(pl.ign_items[pl.ign_ptr[heavy]:pl.ign_ptr[heavy + 1]].tolist())
    pos_hits, neg_hits, rounds = {}, {}, 400
    for _ in range(rounds):
        u, p, n = (x.cpu().numpy() for x in s.sample(256))
        np.add.at(cnt_u, u, 1)
        for j in np.nonzero(u == hu)[0]:
            pos_hits[int(p[j])] = pos_hits.get(int(p[j]), 0) + 1
            neg_hits[int(n[j])] = neg_hits.get(int(n[j]), 0) + 1
    exp = rounds * 256 / n_p
    got = cnt_u[pl.users]
    assert cnt_u.sum() == rounds * 256 and cnt_u[np.setdiff1d(np.arange(g.n_users), pl.users)].sum() == 0
    assert abs(got.mean() - exp) < 1e-9 and got.std() < 3.0 * np.sqrt(exp)       # binomial spread
    assert set(pos_hits) <= set(h_pos.tolist()) and len(set(pos_hits)) >= min(len(set(h_pos.tolist())), 3)
    assert not (set(neg_hits) & h_ign)
