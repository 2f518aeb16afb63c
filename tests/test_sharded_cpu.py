"""Multi-rank host logic of the row-partitioned training step on CPU (`gloo`, world_size 2 and 3;
SURVEY.md 8(e)): partition, source renumbering, all-gather / all-reduce plumbing and the Horner
backward, with a torch restatement of the kernels as the backend (tests/_cpu_backend.py).
The checker is the oracle's single-process reference step (`oracle.port.train_step`)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LR, DECAY, DIM, LAYERS, BATCH = 0.005, 1e-4, 16, 3, 64


def _inputs():
    from gnn_ecommerce_b200 import synth
    from oracle import port
    g = synth.make_graph(600, 90, 4000, seed=21)
    ei, ew = port.df_to_graph(g.user, g.item, g.weight)
    torch.manual_seed(5)
    init = torch.nn.init.xavier_uniform_(torch.empty(g.num_nodes, DIM))
    pl = synth.purchase_lists(g)
    rng = np.random.default_rng(6)
    triples = [tuple(torch.from_numpy(x) for x in synth.sample_triples(pl, BATCH, g.n_users, g.n_items, rng))
               for _ in range(2)]
    return g, ei, ew, init, triples


def _reference():
    from oracle import port
    g, ei, ew, init, triples = _inputs()
    model = port.PortLightGCN(g.num_nodes, DIM, LAYERS)
    with torch.no_grad():
        model.embedding.weight.copy_(init)
    opt = torch.optim.Adam(model.parameters(), LR)
    losses = [port.train_step(model, opt, ei, ew, *t, DECAY) for t in triples]
    with torch.no_grad():
        emb = model.get_embedding(ei, ew)
    return np.array(losses), model.embedding.weight.detach().clone(), emb


def _worker(rank, world, port_no, out, mode="rows"):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _cpu_backend import CpuCheckBackend
    from gnn_ecommerce_b200.sharded import make_sharded_trainer
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port_no}", rank=rank, world_size=world)
    g, ei, ew, init, triples = _inputs()
    tr = make_sharded_trainer(ei, ew, g.num_nodes, DIM, LAYERS, init, mode=mode, lr=LR, backend=CpuCheckBackend(),
                              ld=DIM)
    assert type(tr).__name__ == ("BipartiteShardedTrainer" if mode == "bipartite" else "ShardedBPRTrainer")
    assert sum(tr.part.hi(r) - tr.part.lo(r) for r in range(world)) == (g.n_users if mode == "bipartite"
                                                                        else g.num_nodes)
    losses = [tr.step(*t, DECAY).numpy() for t in triples]
    w, emb = tr.weight(), tr.embedding()
    if rank == 0:
        torch.save({"losses": np.array(losses), "w": w, "emb": emb, "bounds": tr.part.bounds,
                    "local_nnz": tr.local_nnz}, out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(1, "rows"), (2, "rows"), (3, "rows"), (1, "bipartite"), (2, "bipartite"),
                                        (3, "bipartite")])
def test_sharded_step_matches_reference(world, mode, tmp_path):
    out = str(tmp_path / "res.pt")
    port_no = 29500 + (os.getpid() % 2000) + world + (10 if mode == "bipartite" else 0)
    if world == 1:
        _worker(0, 1, port_no, out, mode)
    else:
        mp.spawn(_worker, args=(world, port_no, out, mode), nprocs=world, join=True)
    got = torch.load(out, weights_only=False)
    want_losses, want_w, want_emb = _reference()
    assert np.allclose(got["losses"], want_losses, rtol=1e-5, atol=0)
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
    assert rel(got["w"], want_w) < 5e-4          # two Adam steps amplify last-bit gradient noise
    assert float((got["w"] - want_w).abs().median()) < 1e-7
    assert rel(got["emb"], want_emb) < 5e-4
    b = got["bounds"]
    assert b[0] == 0 and np.all(np.diff(b) >= 0)
    assert b[-1] == (want_w.shape[0] if mode == "rows" else 600)


class _FailingArena:
    """What `CudaBackend.peer_arena` returns, reduced to the two collective steps and their failure modes."""
    closed = 0

    def __init__(self, fail_connect):
        self._fail_connect = fail_connect

    def connect(self):
        dist.barrier()                                   # the all-gather of the handles: every rank must get here
        if self._fail_connect:
            raise RuntimeError("cudaIpcOpenMemHandle: peer access is not supported between these two devices")

    def close(self):
        _FailingArena.closed += 1


def _fallback_worker(rank, world, port_no, out, fail_at):
    import warnings
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _cpu_backend import CpuCheckBackend
    from gnn_ecommerce_b200.sharded import make_sharded_trainer

    class Backend(CpuCheckBackend):
        supports_peer, peer_on_cpu = True, True

        def peer_arena(self, nbytes, device, group):
            if fail_at == "alloc" and rank == 1:
                raise RuntimeError("cudaMalloc: out of memory")
            return _FailingArena(fail_at == "connect" and rank == 0)

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port_no}", rank=rank, world_size=world)
    g, ei, ew, init, triples = _inputs()
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        tr = make_sharded_trainer(ei, ew, g.num_nodes, DIM, LAYERS, init, mode="bipartite", lr=LR, backend=Backend(),
                                  ld=DIM)
    # every rank fell back together (no rank is left waiting in a collective), arenas that did get made are closed
    assert tr.peer is None
    assert any("ncclAllReduce" in str(w.message) for w in caught)
    assert _FailingArena.closed == (0 if (fail_at == "alloc" and rank == 1) else 1)
    losses = [tr.step(*t, DECAY).numpy() for t in triples]
    if rank == 0:
        torch.save({"losses": np.array(losses)}, out)
    # exchange="peer" must raise on every rank instead of falling back
    with pytest.raises(RuntimeError):
        make_sharded_trainer(ei, ew, g.num_nodes, DIM, LAYERS, init, mode="bipartite", lr=LR, backend=Backend(),
                             ld=DIM, exchange="peer")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_at", ["alloc", "connect"])
def test_peer_arena_failure_on_one_rank_falls_back_on_all(fail_at, tmp_path):
    """The peer-memory exchange is set up in two collective phases (`_open_peer_arena`): if one rank cannot
    allocate or map an arena, ALL ranks must agree to use the NCCL path -- none may hang in the all-gather of
    the IPC handles -- and the step must still match the reference."""
    out = str(tmp_path / "res.pt")
    port_no = 31500 + (os.getpid() % 2000) + (1 if fail_at == "connect" else 0)
    mp.spawn(_fallback_worker, args=(2, port_no, out, fail_at), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    want_losses, _, _ = _reference()
    assert np.allclose(got["losses"], want_losses, rtol=1e-5, atol=0)


def test_close_moves_the_item_tables_out_of_the_arena():
    """`close()` frees the peer arena the item tables live in: they must be ordinary tensors afterwards
    (no use after free in `weight()` / `embedding()`), and `step()` must refuse to continue."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _cpu_backend import CpuCheckBackend
    from gnn_ecommerce_b200.sharded import make_sharded_trainer
    g, ei, ew, init, triples = _inputs()
    tr = make_sharded_trainer(ei, ew, g.num_nodes, DIM, LAYERS, init, mode="bipartite", lr=LR, backend=CpuCheckBackend(),
                              ld=DIM)
    tr.step(*triples[0], DECAY)
    w, emb = tr.weight().clone(), tr.embedding().clone()

    class Arena:                                         # stands for the CUDA IPC arena
        closed = False

        def close(self):
            Arena.closed = True
    before = {n: getattr(tr, n) for n in ("e0_i", "m_i", "v_i", "out_i")}
    tr.peer = Arena()
    tr.close()
    assert Arena.closed and tr.peer is None
    for n, t in before.items():
        now = getattr(tr, n)
        assert now is not t and now.data_ptr() != t.data_ptr() and torch.equal(now, t)
    assert torch.equal(tr.weight(), w) and torch.allclose(tr.embedding(), emb)
    with pytest.raises(RuntimeError, match="closed"):
        tr.step(*triples[1], DECAY)


def test_bipartite_split_detection():
    from gnn_ecommerce_b200.sharded import bipartite_split
    _, ei, _, _, _ = _inputs()
    assert bipartite_split(ei) == 600
    assert bipartite_split(torch.tensor([[0, 1, 2], [1, 2, 0]])) is None          # a triangle


def test_row_partition_balances_cost_and_renumbers():
    from gnn_ecommerce_b200.sharded import ROW_COST, RowPartition
    rng = np.random.default_rng(0)
    deg = np.concatenate([rng.integers(1, 6, 5000), rng.integers(50, 4000, 120)])      # users | hub items
    part = RowPartition(deg, 4)
    cost = np.array([(deg[part.lo(r):part.hi(r)] + ROW_COST).sum() for r in range(4)])
    assert cost.max() / cost.mean() < 1.25
    ids = torch.from_numpy(rng.integers(0, len(deg), 1000))
    pid = part.padded_id(ids)
    own = part.owner(ids)
    assert torch.all(own == pid // part.max_rows)
    lo = torch.as_tensor(part.bounds[:-1])[own]
    assert torch.all(pid - own * part.max_rows == ids - lo)
    table = torch.arange(4 * part.max_rows).float()[:, None]
    assert torch.equal(part.unpad(table)[ids, 0], pid.float())


def test_shard_users_covers_everyone():
    from gnn_ecommerce_b200.sharded import shard_users
    for n, w in [(10, 3), (1_600_000, 8), (5, 8)]:
        spans = [shard_users(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
