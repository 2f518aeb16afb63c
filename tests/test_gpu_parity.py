"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors.

Tolerances (BASELINE.json north_star): integer / index work and the fp32 degree
normalisation bit-exact; embeddings, losses and gradients within 1e-5 max-norm relative of the
fp32 reference; post-Adam weights judged against the fp64 reference with the reference's own
fp32-vs-fp64 error as the budget (Adam amplifies last-bit gradient noise, hard part 4).
"""
import os

import numpy as np
import pytest
import torch

from gnn_ecommerce_b200 import synth
from oracle import port
from oracle.lgconv import LGConv as OracleLGConv

pytestmark = pytest.mark.gpu

LR, DECAY = 0.005, 1e-4
DEV = "cuda:0"


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _c1_inputs(z):
    g = synth.make_config_graph("c1")
    dim, layers = int(z["dim"]), int(z["layers"])
    bound = np.sqrt(6.0 / (g.num_nodes + dim))
    init = np.random.default_rng(int(z["init_seed"])).uniform(-bound, bound, (g.num_nodes, dim)).astype(np.float32)
    ei, ew = port.df_to_graph(g.user, g.item, g.weight)
    return g, dim, layers, init, ei, ew


def _model(num_nodes, dim, layers, init):
    from gnn_ecommerce_b200 import LightGCN
    model = LightGCN(num_nodes, dim, layers)
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(init))
    return model.to(DEV)


# --------------------------------------------------------------------------- graph build
def test_graph_build_is_bit_exact(golden_c1):
    from gnn_ecommerce_b200.graph import Graph
    g, _, _, _, ei, ew = _c1_inputs(golden_c1)
    gr = Graph(ei.to(DEV), ew.to(DEV), g.num_nodes)
    a = {k: v.cpu().numpy() for k, v in gr.arrays().items()}
    csr = port.csr_by_target(ei.numpy(), ew.numpy(), g.num_nodes)
    assert np.array_equal(a["rowptr"], csr["rowptr"])
    assert np.array_equal(a["src"], csr["src"])
    assert np.array_equal(a["eid"], csr["eid"])                 # stable: edge order kept per row
    assert np.array_equal(a["deg"], csr["deg"]) and np.array_equal(a["deg"], golden_c1["deg"])
    assert np.array_equal(a["dis"], csr["dis"]) and np.array_equal(a["dis"], golden_c1["dis"])
    assert np.array_equal(a["w_hat"], csr["w_hat_csr"])
    assert np.array_equal(gr.w_hat_edge_order().cpu().numpy(), csr["w_hat_edge"])
    assert gr.is_symmetric
    assert np.array_equal(np.diff(a["rowptr"]), golden_c1["count_deg"])


def test_graph_build_edge_cases():
    from gnn_ecommerce_b200.graph import Graph
    # isolated nodes, duplicate edges, a self loop, no weights, non-symmetric
    ei = torch.tensor([[0, 0, 1, 2, 2, 5], [1, 1, 0, 2, 0, 0]], device=DEV)
    gr = Graph(ei, None, 8)
    a = {k: v.cpu().numpy() for k, v in gr.arrays().items()}
    csr = port.csr_by_target(ei.cpu().numpy(), np.ones(6, np.float32), 8)
    for k in ("rowptr", "src", "eid", "deg", "dis"):
        assert np.array_equal(a[k], csr[k]), k
    assert np.array_equal(a["w_hat"], csr["w_hat_csr"])
    assert not gr.is_symmetric
    assert a["dis"][3] == 0 and a["dis"][7] == 0
    # empty graph
    g0 = Graph(torch.zeros(2, 0, dtype=torch.int64, device=DEV), None, 5)
    assert g0.nnz == 0 and np.array_equal(g0.arrays()["rowptr"].cpu().numpy(), np.zeros(6, np.int32))
    # out-of-range index -> error code, not a device assert
    with pytest.raises(RuntimeError, match="outside"):
        Graph(torch.tensor([[0, 9], [1, 0]], device=DEV), None, 4)


# --------------------------------------------------------------------------- LGConv operator seam
@pytest.mark.parametrize("dim", [16, 64, 80, 90, 128])
def test_lgconv_forward_backward_vs_oracle(dim):
    from gnn_ecommerce_b200 import LGConv
    g = synth.make_graph(3000, 400, 40_000, seed=3)
    ei, ew = port.df_to_graph(g.user, g.item, g.weight)
    x = torch.randn(g.num_nodes, dim, generator=torch.Generator().manual_seed(1))
    xo = x.clone().requires_grad_(True)
    yo = OracleLGConv()(xo, ei, ew)
    go = torch.randn(yo.shape, generator=torch.Generator().manual_seed(2))
    yo.backward(go)
    xg = x.to(DEV).requires_grad_(True)
    yg = LGConv()(xg, ei.to(DEV), ew.to(DEV))
    yg.backward(go.to(DEV))
    assert yg.shape == yo.shape
    assert rel(yg.detach().cpu(), yo.detach()) < 1e-5
    assert rel(xg.grad.cpu(), xo.grad) < 1e-5


def test_lgconv_nonsymmetric_backward_uses_transpose():
    from gnn_ecommerce_b200 import LGConv
    rng = np.random.default_rng(5)
    n, nnz, dim = 500, 6000, 32
    ei = torch.from_numpy(rng.integers(0, n, size=(2, nnz)))
    ew = torch.from_numpy(rng.choice(synth.WEIGHT_VALUES, nnz))
    x = torch.randn(n, dim, generator=torch.Generator().manual_seed(1))
    xo = x.clone().requires_grad_(True)
    yo = OracleLGConv()(xo, ei, ew)
    yo.square().sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    yg = LGConv()(xg, ei.to(DEV), ew.to(DEV))
    yg.square().sum().backward()
    assert rel(yg.detach().cpu(), yo.detach()) < 1e-5
    assert rel(xg.grad.cpu(), xo.grad) < 1e-5


def test_hub_rows_split_across_warps():
    """A star: one item with 5000 in-edges (multi-chunk two-phase path) + light rows."""
    from gnn_ecommerce_b200 import LGConv
    from gnn_ecommerce_b200.graph import graph_for
    n_users = 5000
    u = np.arange(n_users, dtype=np.int64)
    it = np.full(n_users, n_users, dtype=np.int64)
    it[::7] = n_users + 1 + (u[::7] % 40)                      # some medium rows (deg ~ 18)
    w = np.random.default_rng(0).choice(synth.WEIGHT_VALUES, n_users).astype(np.float32)
    ei, ew = port.df_to_graph(u, it, w)
    n = n_users + 41
    x = torch.randn(n, 64, generator=torch.Generator().manual_seed(1))
    yo = OracleLGConv()(x, ei, ew)
    eig, ewg = ei.to(DEV), ew.to(DEV)
    yg = LGConv()(x.to(DEV), eig, ewg)
    info = graph_for(eig, ewg, n).info
    assert info.num_split_rows >= 1 and info.num_chunks > info.num_split_rows
    assert rel(yg.cpu(), yo) < 1e-5


# --------------------------------------------------------------------------- get_embedding / forward
def test_get_embedding_matches_golden_c1(golden_c1):
    z = golden_c1
    g, dim, layers, init, ei, ew = _c1_inputs(z)
    model = _model(g.num_nodes, dim, layers, init)
    with torch.no_grad():
        out = model.get_embedding(ei.to(DEV), ew.to(DEV)).cpu().numpy()
    assert out.shape == (g.num_nodes, dim)
    assert rel(out[z["rows"]], z["f32_out0_rows"]) < 1e-5
    assert rel(out[z["rows"]], z["f64_out0_rows"]) < 1e-5


@pytest.mark.parametrize("tag", ["", "iso_"])
def test_get_embedding_matches_golden_tiny(golden_tiny, tag):
    z = golden_tiny
    n = int(z["num_nodes_iso"]) if tag else int(z["n_users"]) + int(z["n_items"])
    init = z["init_iso"] if tag else z["init"]
    model = _model(n, int(z["dim"]), int(z["layers"]), init)
    ei, ew = torch.from_numpy(z["edge_index"]).to(DEV), torch.from_numpy(z["edge_weight"]).to(DEV)
    with torch.no_grad():
        out = model.get_embedding(ei, ew).cpu().numpy()
    assert rel(out, z[f"f32_{tag}out0"]) < 1e-5
    if tag:   # isolated nodes: alpha_0 * E0 exactly, nothing propagated into them
        assert np.array_equal(out[n - 3:], z["f32_iso_out0"][n - 3:])


@pytest.mark.parametrize("dim,layers", [(64, 3), (80, 4), (90, 5), (64, 1), (32, 0)])
def test_forward_and_autograd_vs_oracle(dim, layers):
    """README grid of the reference: K in {3,4,5}, d in {64,80,90} (src README.md:61-64)."""
    g = synth.make_graph(4000, 600, 50_000, seed=9)
    ei, ew = port.df_to_graph(g.user, g.item, g.weight)
    torch.manual_seed(3)
    om = port.PortLightGCN(g.num_nodes, dim, layers)
    model = _model(g.num_nodes, dim, layers, om.embedding.weight.detach().numpy())
    pl = synth.purchase_lists(g)
    u, p, n = (torch.from_numpy(x) for x in synth.sample_triples(pl, 256, g.n_users, g.n_items,
                                                                np.random.default_rng(4)))
    labels = port.batch_pos_neg_edges(u, p, n)
    so = om(ei, labels, ew)
    lo = port.bpr_loss(so[:256], so[256:]) + port.regularization_loss(om.embedding.weight, 256, u, p, n, DECAY)
    lo.backward()
    sg = model(ei.to(DEV), labels.to(DEV), ew.to(DEV))
    lg = model.recommendation_loss(sg[:256], sg[256:], 0) * 256 + port.regularization_loss(
        model.embedding.weight, 256, u.to(DEV), p.to(DEV), n.to(DEV), DECAY)
    lg.backward()
    assert rel(sg.detach().cpu(), so.detach()) < 1e-5
    assert abs(lg.item() - lo.item()) <= 1e-5 * abs(lo.item())
    assert rel(model.embedding.weight.grad.cpu(), om.embedding.weight.grad) < 1e-5


# --------------------------------------------------------------------------- training step
def _budget(new, f32, f64, floor=1e-6):
    """err(new, fp64) <= 1.5 * err(reference fp32, fp64) + floor (max-norm relative). The measured
    ratio of every call is appended to gpurun_out/adam_budget.log (evidence for the 1.5)."""
    mine, ref = rel(new, f64), rel(f32, f64)
    try:
        import inspect
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "adam_budget.log"), "a") as f:
            f.write(f"{inspect.stack()[1].function}: err(new, fp64) = {mine:.3e}, err(reference fp32, fp64) = {ref:.3e}, "
                    f"ratio = {mine / max(ref, 1e-30):.3f}\n")
    except OSError:
        pass
    return mine <= 1.5 * ref + floor


def test_autograd_training_matches_golden_c1(golden_c1):
    """Drop-in usage: reference loop with torch.optim.Adam, our module underneath."""
    z = golden_c1
    g, dim, layers, init, ei, ew = _c1_inputs(z)
    model = _model(g.num_nodes, dim, layers, init)
    opt = torch.optim.Adam(model.parameters(), LR)
    eig, ewg = ei.to(DEV), ew.to(DEV)
    rows = z["rows"]
    for s, t in enumerate(z["triples"]):
        u, p, n = (torch.from_numpy(np.ascontiguousarray(x)).to(DEV) for x in t)
        opt.zero_grad()
        out = model(eig, port.batch_pos_neg_edges(u, p, n), ewg)
        size = len(u)
        bpr = model.recommendation_loss(out[:size], out[size:], 0) * size
        reg = port.regularization_loss(model.embedding.weight, size, u, p, n, DECAY)
        loss = bpr + reg
        loss.backward()
        if s == 0:
            assert rel(out.detach().cpu(), z["f32_scores0"]) < 1e-5
            assert rel(model.embedding.weight.grad.cpu().numpy()[rows], z["f32_grad0_rows"]) < 1e-5
        opt.step()
        got = np.array([bpr.item(), reg.item(), loss.item()])
        assert np.allclose(got, z["f32_losses"][s], rtol=1e-5, atol=0)
        w = model.embedding.weight.detach().cpu().numpy()[rows]
        assert _budget(w, z[f"f32_w{s + 1}_rows"], z[f"f64_w{s + 1}_rows"])


def test_fused_step_matches_golden_c1(golden_c1):
    from gnn_ecommerce_b200 import FusedBPRTrainer
    z = golden_c1
    g, dim, layers, init, ei, ew = _c1_inputs(z)
    model = _model(g.num_nodes, dim, layers, init)
    trainer = FusedBPRTrainer(model, lr=LR)
    eig, ewg = ei.to(DEV), ew.to(DEV)
    rows = z["rows"]
    for s, t in enumerate(z["triples"]):
        u, p, n = (torch.from_numpy(np.ascontiguousarray(x)).to(DEV) for x in t)
        loss3 = trainer.step(eig, ewg, u, p, n, DECAY).cpu().numpy()
        assert np.allclose(loss3, z["f32_losses"][s], rtol=1e-5, atol=0)
        w = model.embedding.weight.detach().cpu().numpy()[rows]
        assert _budget(w, z[f"f32_w{s + 1}_rows"], z[f"f64_w{s + 1}_rows"])
    sd = trainer.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and int(sd["state"][0]["step"]) == 2
    assert list(model.state_dict().keys()) == ["alpha", "embedding.weight"]
    assert model.state_dict()["embedding.weight"].shape == (g.num_nodes, dim)


@pytest.mark.parametrize("tag", ["", "iso_"])
def test_fused_step_matches_golden_tiny(golden_tiny, tag):
    """Duplicate triples (gradients accumulate, L2 counts multiplicity) and isolated nodes."""
    from gnn_ecommerce_b200 import FusedBPRTrainer
    z = golden_tiny
    n = int(z["num_nodes_iso"]) if tag else int(z["n_users"]) + int(z["n_items"])
    model = _model(n, int(z["dim"]), int(z["layers"]), z["init_iso"] if tag else z["init"])
    trainer = FusedBPRTrainer(model, lr=LR)
    ei, ew = torch.from_numpy(z["edge_index"]).to(DEV), torch.from_numpy(z["edge_weight"]).to(DEV)
    for s, t in enumerate(z["triples"]):
        u, p, nn_ = (torch.from_numpy(np.ascontiguousarray(x)).to(DEV) for x in t)
        loss3 = trainer.step(ei, ew, u, p, nn_, DECAY).cpu().numpy()
        assert np.allclose(loss3, z[f"f32_{tag}losses"][s], rtol=1e-5, atol=0)
    w = model.embedding.weight.detach().cpu().numpy()
    assert _budget(w, z[f"f32_{tag}w2"], z[f"f64_{tag}w2"], floor=2e-6)


@pytest.mark.parametrize("dim,layers", [(90, 5), (80, 4), (64, 1), (16, 0)])
def test_fused_step_equals_autograd_path(dim, layers):
    """Padded storage (d=90 -> 96 floats per row), other depths: fused == autograd + torch Adam."""
    from gnn_ecommerce_b200 import FusedBPRTrainer
    g = synth.make_graph(3000, 500, 40_000, seed=13)
    ei, ew = port.df_to_graph(g.user, g.item, g.weight)
    eig, ewg = ei.to(DEV), ew.to(DEV)
    torch.manual_seed(7)
    init = torch.nn.init.xavier_uniform_(torch.empty(g.num_nodes, dim)).numpy()
    ma, mf = _model(g.num_nodes, dim, layers, init), _model(g.num_nodes, dim, layers, init)
    opt, trainer = torch.optim.Adam(ma.parameters(), LR), FusedBPRTrainer(mf, lr=LR)
    pl = synth.purchase_lists(g)
    rng = np.random.default_rng(2)
    for _ in range(3):
        u, p, n = (torch.from_numpy(x).to(DEV) for x in synth.sample_triples(pl, 200, g.n_users, g.n_items, rng))
        opt.zero_grad()
        out = ma(eig, port.batch_pos_neg_edges(u, p, n), ewg)
        bpr = ma.recommendation_loss(out[:200], out[200:], 0) * 200
        reg = port.regularization_loss(ma.embedding.weight, 200, u, p, n, DECAY)
        (bpr + reg).backward()
        opt.step()
        loss3 = trainer.step(eig, ewg, u, p, n, DECAY)
        assert np.allclose(loss3.cpu().numpy(), [bpr.item(), reg.item(), (bpr + reg).item()], rtol=2e-5)
    wa, wf = ma.embedding.weight.detach().cpu(), mf.embedding.weight.detach().cpu()
    assert wf.shape == (g.num_nodes, dim)
    assert rel(wf, wa) < 5e-4           # three Adam steps amplify last-bit gradient differences
    assert np.median(np.abs(wf.numpy() - wa.numpy())) < 1e-7


def test_adam_kernel_in_isolation():
    """Identical gradients in, torch.optim.Adam (CPU) as the checker: <= 1e-6 (hard part 4a)."""
    from gnn_ecommerce_b200 import ops
    gen = torch.Generator().manual_seed(0)
    p = torch.randn(1000, 64, generator=gen) * 1e-2
    po = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([po], LR)
    pg, m, v = p.to(DEV), torch.zeros(1000, 64, device=DEV), torch.zeros(1000, 64, device=DEV)
    for step in range(1, 6):
        grad = torch.randn(1000, 64, generator=gen) * (1e-6 if step % 2 else 1e-3)
        grad[::5] = 0
        po.grad = grad.clone()
        opt.step()
        ops.adam_step(pg, grad.to(DEV), m, v, LR, step=step)
        assert rel(pg.cpu(), po.detach()) < 1e-6
    st = opt.state[po]
    assert rel(m.cpu(), st["exp_avg"]) < 1e-6 and rel(v.cpu(), st["exp_avg_sq"]) < 1e-6


# --------------------------------------------------------------------------- full size (c2)
def test_c2_full_size_against_reference_gpu_path():
    """At BASELINE.json's full size the oracle's LGConv restatement runs on the GPU itself
    (index_select / mul / scatter_add_ = the reference's own CUDA path), plus size-independent
    properties: linearity and symmetry <y, A x> = <A y, x>."""
    from gnn_ecommerce_b200 import LGConv
    g = synth.make_config_graph("c2")
    ei, ew = port.df_to_graph(g.user, g.item, g.weight)
    ei, ew = ei.to(DEV), ew.to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(g.num_nodes, 64, device=DEV, generator=gen)
    y = torch.randn(g.num_nodes, 64, device=DEV, generator=gen)
    conv = LGConv()
    ax = conv(x, ei, ew)
    ref = OracleLGConv()(x, ei, ew)
    assert rel(ax.cpu(), ref.cpu()) < 1e-5
    del ref
    ay = conv(y, ei, ew)
    lhs, rhs = (y.double() * ax.double()).sum().item(), (ay.double() * x.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs))
    comb = conv(2.0 * x - 0.5 * y, ei, ew)
    assert rel(comb.cpu(), (2.0 * ax - 0.5 * ay).cpu()) < 1e-5


# --------------------------------------------------------------------------- sharded driver, 1 rank
@pytest.mark.parametrize("mode", ["rows", "bipartite"])
@pytest.mark.parametrize("dim,layers", [(64, 3), (90, 5), (16, 1)])
def test_sharded_trainer_single_rank_equals_fused(dim, layers, mode):
    """The row-partitioned driver (lgc_graph_build_rect + lgc_spmm_ex + compact BPR rows) with one
    rank must reproduce the fused single-GPU step; multi-rank orchestration is covered on CPU
    (tests/test_sharded_cpu.py) and on 2+ GPUs by bench.py."""
    from gnn_ecommerce_b200 import FusedBPRTrainer
    from gnn_ecommerce_b200.sharded import make_sharded_trainer
    g = synth.make_graph(3000, 500, 40_000, seed=13)
    ei, ew = port.df_to_graph(g.user, g.item, g.weight)
    eig, ewg = ei.to(DEV), ew.to(DEV)
    torch.manual_seed(7)
    init = torch.nn.init.xavier_uniform_(torch.empty(g.num_nodes, dim))
    mf = _model(g.num_nodes, dim, layers, init.numpy())
    fused = FusedBPRTrainer(mf, lr=LR)
    sharded = make_sharded_trainer(eig, ewg, g.num_nodes, dim, layers, init, mode=mode, lr=LR)
    pl = synth.purchase_lists(g)
    rng = np.random.default_rng(2)
    for _ in range(5):                 # steps 1-2 eager, step 3 captures a CUDA graph, 4-5 replay it
        u, p, n = (torch.from_numpy(x).to(DEV) for x in synth.sample_triples(pl, 200, g.n_users, g.n_items, rng))
        a = fused.step(eig, ewg, u, p, n, DECAY).cpu().numpy()
        b = sharded.step(u, p, n, DECAY).cpu().numpy()
        assert np.allclose(a, b, rtol=2e-5)
    assert sharded._graph is not None
    wf, ws = mf.embedding.weight.detach().cpu(), sharded.weight().cpu()
    assert ws.shape == (g.num_nodes, dim)
    assert rel(ws, wf) < 1e-3 and np.median(np.abs(ws.numpy() - wf.numpy())) < 1e-7
    with torch.no_grad():
        ef = mf.get_embedding(eig, ewg).cpu()
    assert rel(sharded.embedding().cpu(), ef) < 5e-4


# --------------------------------------------------------------------------- checkpoint wire format
def test_checkpoint_roundtrip_in_the_reference_format(tmp_path):
    """`save_model` writes the reference's dict (src/utils_v2.py:212-230) with a LOGICAL [N, d]
    table although d = 90 is stored padded; resuming from it continues like the uninterrupted run
    (to rounding), and torch.optim.Adam accepts the optimizer state."""
    from gnn_ecommerce_b200 import FusedBPRTrainer, load_model, save_model
    g = synth.make_graph(2000, 300, 20_000, seed=17)
    ei, ew = port.df_to_graph(g.user, g.item, g.weight)
    eig, ewg = ei.to(DEV), ew.to(DEV)
    torch.manual_seed(1)
    init = torch.nn.init.xavier_uniform_(torch.empty(g.num_nodes, 90)).numpy()
    model = _model(g.num_nodes, 90, 3, init)
    trainer = FusedBPRTrainer(model, lr=LR)
    pl = synth.purchase_lists(g)
    rng = np.random.default_rng(8)
    batches = [tuple(torch.from_numpy(x).to(DEV) for x in synth.sample_triples(pl, 128, g.n_users, g.n_items, rng))
               for _ in range(4)]
    for b in batches[:2]:
        trainer.step(eig, ewg, *b, DECAY)
    path = str(tmp_path / "LightGCN_best.pt")
    save_model(path, model, trainer, 0.1, 0.2, epoch=3, hyperparams={"latent_dim": 90, "n_layers": 3})
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"timestamp", "epoch", "model_state_dict", "optimizer_state_dict", "precision", "recall",
                       "hyperparams"}
    assert list(ck["model_state_dict"]) == ["alpha", "embedding.weight"]
    assert ck["model_state_dict"]["embedding.weight"].shape == (g.num_nodes, 90)
    ref_param = torch.nn.Parameter(ck["model_state_dict"]["embedding.weight"].clone())
    torch.optim.Adam([ref_param], LR).load_state_dict(ck["optimizer_state_dict"])      # torch accepts it
    model2, trainer2, _ = load_model(path, DEV)
    assert trainer2.step_count == 2
    for b in batches[2:]:
        a = trainer.step(eig, ewg, *b, DECAY).cpu()
        c = trainer2.step(eig, ewg, *b, DECAY).cpu()
        # duplicates inside a batch meet in float atomics, so two runs agree to rounding, not bit for bit;
        # a resume that lost the moments or the step count would be off by ~lr per weight.
        assert torch.allclose(a, c, rtol=1e-5, atol=1e-9), (a, c)
    dw = (model.embedding.weight.detach() - model2.embedding.weight.detach()).abs().max().item()
    assert dw < 1e-6, dw


# --------------------------------------------------------------------------- graph ingest from pairs
def test_graph_from_interaction_pairs_is_bit_identical_to_df_to_graph():
    """`Graph.from_interactions` (lgc_graph_build_pairs) == `Graph` on df_to_graph's COO
    (reference src/utils_v2.py:146-165): same CSR, same normalised weights, and the module accepts the
    prebuilt graph where it takes edge_index."""
    from gnn_ecommerce_b200.graph import Graph
    g = synth.make_graph(3000, 200, 20_000, seed=23)
    ei, ew = port.df_to_graph(g.user, g.item, g.weight)
    ref = Graph(ei.to(DEV), ew.to(DEV), g.num_nodes)
    got = Graph.from_interactions(torch.from_numpy(g.user).to(DEV), torch.from_numpy(g.item).to(DEV),
                                  torch.from_numpy(g.weight).to(DEV), g.num_nodes)
    assert got.nnz == ref.nnz and got.is_symmetric
    a, b = ref.arrays(), got.arrays()
    for k in ("rowptr", "src", "eid", "w_hat", "deg", "dis"):
        assert torch.equal(a[k], b[k]), k
    torch.manual_seed(3)
    init = torch.nn.init.xavier_uniform_(torch.empty(g.num_nodes, 64)).numpy()
    model = _model(g.num_nodes, 64, 3, init)
    with torch.no_grad():
        e_ref = model.get_embedding(ei.to(DEV), ew.to(DEV))
        e_got = model.get_embedding(got, None)
    assert torch.equal(e_ref, e_got)



def test_resume_from_a_checkpoint_written_by_the_reference():
    """tests/golden/ref_checkpoint_tiny.pt was written by the reference's own `save_model` after two
    reference steps (make_ref_checkpoint.py); `load_model` restores weights, Adam moments and the step
    count, and the next fused step lands where the reference's third step landed."""
    import os
    from gnn_ecommerce_b200 import load_model
    here = os.path.dirname(os.path.abspath(__file__))
    z = np.load(os.path.join(here, "golden", "tiny.npz"))
    nxt = np.load(os.path.join(here, "golden", "ref_checkpoint_tiny_next.npz"))
    model, trainer, ck = load_model(os.path.join(here, "golden", "ref_checkpoint_tiny.pt"), DEV)
    assert trainer.step_count == 2 and trainer.lr == ck["hyperparams"]["lr"]
    ei, ew = torch.from_numpy(z["edge_index"]).to(DEV), torch.from_numpy(z["edge_weight"]).to(DEV)
    u, p, n = (torch.from_numpy(np.ascontiguousarray(x)).to(DEV) for x in nxt["triple"])
    loss3 = trainer.step(ei, ew, u, p, n, DECAY).cpu().numpy()
    assert np.allclose(loss3, nxt["losses"], rtol=1e-5, atol=0), (loss3, nxt["losses"])
    assert rel(model.embedding.weight.detach().cpu().numpy(), nxt["w3"]) < 1e-5


def test_rectangular_operator_entirely_in_the_sweep_with_empty_rows():
    """The item-partial operator of the multi-GPU step: a small rectangular operator goes to the sweep
    kernel as a whole, including rows WITHOUT local edges (an item none of this rank's users touched),
    which must come out as exact zeros / as the pure epilogue."""
    from gnn_ecommerce_b200.sharded import CudaBackend
    be = CudaBackend()
    rng = np.random.default_rng(8)
    n_rows, n_cols, nnz, ld = 3000, 50_000, 40_000, 64
    dst = torch.from_numpy(rng.integers(0, n_rows // 2, nnz) * 2)            # odd rows have no edges at all
    src = torch.from_numpy(rng.integers(0, n_cols, nnz))
    w = torch.from_numpy(rng.random(nnz).astype(np.float32))
    h = be.build_rect(src.to(DEV), dst.to(DEV), w.to(DEV), n_rows, n_cols)
    x = torch.randn(n_cols, ld, device=DEV)
    add = torch.randn(n_rows, ld, device=DEV)
    y = torch.full((n_rows, ld), float("nan"), device=DEV)
    ws = be.workspace(h, ld, DEV)
    be.spmm_ex(h, ld, x, ws, 0, y=y, addend=add, scale=2.0, beta=0.5)
    want = torch.zeros(n_rows, ld, dtype=torch.float64).index_add_(0, dst, w.double()[:, None] * x.cpu().double()[src])
    want = 2.0 * want + 0.5 * add.cpu().double()
    assert not torch.isnan(y).any()
    assert rel(y.cpu(), want) < 1e-5
    assert torch.equal(y[1::2], (0.5 * add[1::2]))
    be.destroy(h)
