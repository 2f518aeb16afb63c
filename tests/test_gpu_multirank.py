"""Two ranks over NCCL on two GPUs (skipped on a one-GPU box): the sharded training steps must
reproduce the single-GPU fused step. The host-side sharding logic is covered on CPU under gloo
(tests/test_sharded_cpu.py); this is the NCCL + CUDA-graph path itself.

    gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu -q
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LR, DECAY = 0.005, 1e-4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, mode, dim, layers, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from gnn_ecommerce_b200 import synth
    from gnn_ecommerce_b200.sharded import make_sharded_trainer
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    g = synth.make_graph(6000, 900, 80_000, seed=21)
    ei = torch.from_numpy(g.edge_index()).to(dev)
    ew = torch.from_numpy(g.edge_weight()).to(dev)
    torch.manual_seed(3)
    init = torch.nn.init.xavier_uniform_(torch.empty(g.num_nodes, dim))
    pl = synth.purchase_lists(g)
    results = {}
    # bipartite: the item rows go through the peer-memory kernel (lgc_item_exchange) AND, as a second
    # trainer, through ncclAllReduce + lgc_epilogue_apply; with two ranks a + b has one rounding, so the
    # two must agree bit for bit
    for exch in (("nccl", "peer") if mode == "bipartite" else (None,)):
        kw = {"exchange": exch} if exch else {}
        tr = make_sharded_trainer(ei, ew, g.num_nodes, dim, layers, init, mode=mode, lr=LR, **kw)
        if exch == "peer":
            assert tr.peer is not None and tr.peer.world == world
        rng = np.random.default_rng(9)
        losses = []
        for _ in range(5):                       # 2 eager steps, 1 capture, 2 CUDA-graph replays
            u, p, n = (torch.from_numpy(x).to(dev) for x in synth.sample_triples(pl, 256, g.n_users, g.n_items, rng))
            losses.append(tr.step(u, p, n, DECAY).cpu().numpy())
        w, emb = tr.weight().cpu().numpy(), tr.embedding().cpu().numpy()
        if hasattr(tr, "check_exchange"):
            tr.check_exchange()
            if exch == "peer":
                assert tr.peer.status()[1] == 5 * 2 * layers + layers      # 5 steps + the embedding() call
        results[exch] = (np.array(losses), w, emb)
        if hasattr(tr, "close"):
            tr.close()
        elif hasattr(tr, "release_graph"):
            tr.release_graph()
        del tr
    if mode == "bipartite":
        for a, b in zip(results["nccl"], results["peer"]):
            assert np.array_equal(a, b), "peer-memory exchange differs from ncclAllReduce + epilogue"
    losses, w, emb = results["peer" if mode == "bipartite" else None]
    if rank == 0:
        np.savez(os.path.join(out_dir, f"{mode}.npz"), losses=losses, w=w, emb=emb)
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["bipartite", "rows"])
@pytest.mark.parametrize("dim,layers", [(64, 3), (90, 5)])
def test_two_rank_nccl_step_equals_single_gpu(tmp_path, mode, dim, layers):
    import torch.multiprocessing as mp
    from gnn_ecommerce_b200 import FusedBPRTrainer, LightGCN, synth
    port = _free_port()
    mp.spawn(_worker, args=(2, port, mode, dim, layers, str(tmp_path)), nprocs=2, join=True)
    z = np.load(os.path.join(str(tmp_path), f"{mode}.npz"))
    dev = "cuda:0"
    g = synth.make_graph(6000, 900, 80_000, seed=21)
    ei = torch.from_numpy(g.edge_index()).to(dev)
    ew = torch.from_numpy(g.edge_weight()).to(dev)
    torch.manual_seed(3)
    init = torch.nn.init.xavier_uniform_(torch.empty(g.num_nodes, dim))
    model = LightGCN(g.num_nodes, dim, layers)
    with torch.no_grad():
        model.embedding.weight.copy_(init)
    model = model.to(dev)
    tr = FusedBPRTrainer(model, lr=LR)
    pl = synth.purchase_lists(g)
    rng = np.random.default_rng(9)
    for s in range(5):
        u, p, n = (torch.from_numpy(x).to(dev) for x in synth.sample_triples(pl, 256, g.n_users, g.n_items, rng))
        want = tr.step(ei, ew, u, p, n, DECAY).cpu().numpy()
        assert np.allclose(z["losses"][s], want, rtol=2e-5), (s, z["losses"][s], want)
    w = model.embedding.weight.detach().cpu().numpy()
    assert np.abs(z["w"] - w).max() / np.abs(w).max() < 1e-3 and np.median(np.abs(z["w"] - w)) < 1e-7
    with torch.no_grad():
        emb = model.get_embedding(ei, ew).cpu().numpy()
    assert np.abs(z["emb"] - emb).max() / np.abs(emb).max() < 5e-4
