"""Full-size parity of the fused training step (BASELINE.json configs c2 and c3): one step of
`lgc_train_step` on the GPU against ONE step of the CPU reference (fp32) and of the fp64 oracle on
the same graph, initial table and triples. 5-25 s of CPU work per case on the GPU box.

Bars: forward embeddings within 1e-5 (max-norm relative) of the fp64 oracle and within 1e-5 + the
reference's own fp32 error of the fp32 reference; losses within 1e-5 of the fp32 reference; the
post-Adam weights within the reference's own fp32-vs-fp64 error of the fp64 result (Adam amplifies
last-bit gradient noise: SURVEY.md 7, hard part 4)."""
import numpy as np
import pytest
import torch

from gnn_ecommerce_b200 import synth
from oracle import port

pytestmark = pytest.mark.gpu

LR, DECAY, BATCH = 0.005, 1e-4, 1024
DEV = "cuda:0"


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("config", ["c2", "c3"])
def test_full_size_fused_step_vs_cpu_reference(config):
    from gnn_ecommerce_b200 import FusedBPRTrainer, LightGCN
    n_users, n_items, n_edges, dim, layers = synth.CONFIGS[config]
    g = synth.make_graph(n_users, n_items, n_edges, seed=42)
    ei, ew = port.df_to_graph(g.user, g.item, g.weight)
    bound = np.sqrt(6.0 / (g.num_nodes + dim))
    init = np.random.default_rng(43).uniform(-bound, bound, (g.num_nodes, dim)).astype(np.float32)
    pl = synth.purchase_lists(g)
    u, p, n = (torch.from_numpy(x) for x in synth.sample_triples(pl, BATCH, g.n_users, g.n_items,
                                                                np.random.default_rng(44)))
    torch.set_num_threads(torch.get_num_threads())

    def cpu(dtype):
        m = port.PortLightGCN(g.num_nodes, dim, layers, dtype=dtype)
        with torch.no_grad():
            m.embedding.weight.copy_(torch.from_numpy(init))
            out = m.get_embedding(ei, ew).numpy()
        opt = torch.optim.Adam(m.parameters(), LR)
        losses = port.train_step(m, opt, ei, ew, u, p, n, DECAY)
        return out, losses, m.embedding.weight.detach().numpy()

    out32, loss32, w32 = cpu(torch.float32)
    out64, loss64, w64 = cpu(torch.float64)

    model = LightGCN(g.num_nodes, dim, layers)
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(init))
    model = model.to(DEV)
    eig, ewg = ei.to(DEV), ew.to(DEV)
    with torch.no_grad():
        out = model.get_embedding(eig, ewg).cpu().numpy()
    # The hub rows sum up to 76 K terms: the reference's own sequential fp32 sum is ~1e-5 away from
    # the exactly-summed result there. Yardstick: the same propagation with the SAME fp32 normalised
    # weights (bit-exact to gcn_norm, tests/test_gpu_parity.py) accumulated in fp64. Bars: this
    # implementation within 1e-5 of that; and of the fp32 reference within 1e-5 + the reference's own
    # summation error against it.
    graph = model.graph(eig, ewg)
    w_hat = graph.w_hat_edge_order().double()
    x = torch.from_numpy(init).to(DEV).double()
    exact = x * float(model.alpha[0])
    for l in range(layers):
        x = torch.zeros_like(x).index_add_(0, eig[1], w_hat.view(-1, 1) * x[eig[0]])
        exact = exact + x * float(model.alpha[l + 1])
    exact = exact.cpu().numpy()
    del x, w_hat
    e_out, e_ref_out = rel(out, exact), rel(out32, exact)
    print(f"{config}: embeddings err(new, fp64 sums) = {e_out:.3e}, err(reference fp32, fp64 sums) = {e_ref_out:.3e}, "
          f"new vs fp32 reference {rel(out, out32):.3e}, fp64 oracle incl. fp64 weights {rel(out32, out64):.3e}")
    assert e_out < 1e-5
    assert rel(out, out32) < 1e-5 + e_ref_out
    trainer = FusedBPRTrainer(model, lr=LR)
    got = trainer.step(eig, ewg, u.to(DEV), p.to(DEV), n.to(DEV), DECAY).cpu().numpy()
    assert np.allclose(got, loss32, rtol=1e-5, atol=0), (got, loss32)
    w = model.embedding.weight.detach().cpu().numpy()
    e_new, e_ref = rel(w, w64), rel(w32, w64)
    m_new, m_ref = float(np.abs(w - w64).mean()), float(np.abs(w32 - w64).mean())
    print(f"{config}: post-Adam max-norm err(new, fp64) = {e_new:.3e}, err(reference fp32, fp64) = {e_ref:.3e}, "
          f"ratio {e_new / e_ref:.2f}; mean |err| new {m_new:.3e}, reference {m_ref:.3e}, ratio {m_new / m_ref:.3f}")
    # Adam turns a gradient entry of ~1e-8 into a step of up to lr: the max over 1e8 weights is set by a
    # handful of such entries in either implementation (measured ratio 1.05 at c2), so the max-norm
    # bar keeps a 1.25 margin and the MEAN error carries the 1.0x bar (+2 %).
    assert e_new <= 1.25 * e_ref + 1e-6, (e_new, e_ref)
    assert m_new <= 1.02 * m_ref + 1e-12, (m_new, m_ref)
    # every weight moved by at most lr (Adam's first step is +-lr * g / (|g| + eps))
    assert np.abs(w - init).max() <= LR * (1 + 1e-4)
    # the step is run-to-run deterministic (no atomics anywhere in it)
    model2 = LightGCN(g.num_nodes, dim, layers)
    with torch.no_grad():
        model2.embedding.weight.copy_(torch.from_numpy(init))
    model2 = model2.to(DEV)
    got2 = FusedBPRTrainer(model2, lr=LR).step(eig, ewg, u.to(DEV), p.to(DEV), n.to(DEV), DECAY).cpu().numpy()
    assert np.array_equal(got, got2)
    assert torch.equal(model.embedding.weight.detach(), model2.embedding.weight.detach())
