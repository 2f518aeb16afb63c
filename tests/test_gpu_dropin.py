"""Drop-in proof at the reference's own call sites: the reference's
`TrainLightGCN` (src/train_lightgcn.py:8-162), `InferenceLightGCN` (src/inference_lightgcn.py:15-48)
and TorchServe `LightGCNHandler` (torchserve/lightgcn_handler.py:9-110) are imported UNMODIFIED and
run with `from lightgcn import LightGCN` resolving to this repo's module; the same code is run with
the reference's own `LightGCN` (over the LGConv restatement) and the results are compared.

The reference sources come from `/root/reference` or the staged copy `oracle/_ref`
(`oracle/stage_reference.py`); without either the tests skip."""
import importlib
import os
import random
import sys
import types

import numpy as np
import pandas as pd
import pytest
import torch

from gnn_ecommerce_b200 import synth
from oracle import reference_shim

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference_shim.reference_available(), reason="reference sources not staged")]

DEV = "cuda:0"
DIM, LAYERS, K = 64, 3, 20


def _raw_frame(seed=3):
    """A c1-sized interaction frame with RAW ids (strings / sparse ints), as the reference reads it
    from its preprocessed CSV (user_id, item_id, weight)."""
    g = synth.make_graph(4000, 600, 40_000, seed=seed)
    rng = np.random.default_rng(seed)
    user_names = rng.permutation(10_000_000)[:g.n_users] + 1000          # sparse raw ids
    item_names = rng.permutation(900_000)[:g.n_items] + 5
    w = g.weight.astype(np.float64).copy()
    return pd.DataFrame({"user_id": user_names[g.user], "item_id": item_names[g.item - g.n_users], "weight": w})


def _dropin_module():
    import gnn_ecommerce_b200.lightgcn as ours
    return ours


def _trainers(tmp_path):
    csv = os.path.join(tmp_path, "interactions.csv")
    _raw_frame().to_csv(csv, index=False)
    out = []
    for tag, mod in (("ours", _dropin_module()), ("ref", None)):
        tl = reference_shim.load_reference_trainer(mod)
        ckpt = os.path.join(tmp_path, tag) + "/"
        os.makedirs(ckpt, exist_ok=True)
        np.random.seed(0)                      # train_test_split draws from numpy's global RNG
        out.append((tl, tl.TrainLightGCN(csv, ckpt, gpu=0), ckpt))
    return out


def test_reference_training_and_eval_loop_over_the_dropin_module(tmp_path):
    (tl_o, t_o, ck_o), (tl_r, t_r, ck_r) = _trainers(str(tmp_path))
    assert tl_o.LightGCN is _dropin_module().LightGCN and tl_r.LightGCN is not tl_o.LightGCN
    assert (t_o.n_users, t_o.n_items, t_o.train_size) == (t_r.n_users, t_r.n_items, t_r.train_size)
    assert torch.equal(t_o.edge_index, t_r.edge_index) and t_o.edge_index.is_cuda
    n = t_o.n_users + t_o.n_items
    torch.manual_seed(11)
    m_r = tl_r.LightGCN(n, DIM, LAYERS)
    m_o = tl_o.LightGCN(n, DIM, LAYERS)
    m_o.load_state_dict(m_r.state_dict())
    m_r.to(t_r.device); m_o.to(t_o.device)
    opt_r = torch.optim.Adam(m_r.parameters(), 0.005)
    opt_o = torch.optim.Adam(m_o.parameters(), 0.005)
    # mini_batch_loop (src/train_lightgcn.py:123-153): batch_loader draws with the `random` module
    random.seed(5); loss_r = t_r.mini_batch_loop(m_r, opt_r, 256, 1e-4, 4)
    random.seed(5); loss_o = t_o.mini_batch_loop(m_o, opt_o, 256, 1e-4, 4)
    assert np.allclose(loss_o, loss_r, rtol=2e-5), (loss_o, loss_r)
    w_o, w_r = m_o.embedding.weight.detach().cpu(), m_r.embedding.weight.detach().cpu()
    assert float((w_o - w_r).abs().max() / w_r.abs().max()) < 2e-3      # 4 Adam steps: see hard part 4
    assert float((w_o - w_r).abs().median()) < 1e-7
    # test() (:155-162) = recommendK + MARK_MAPK on the validation users, same weights in both models
    m_o.load_state_dict(m_r.state_dict())
    p_r, r_r, met_r = t_r.test(m_r, t_r.val_pos_list_df, t_r.val_interactions_t, K)
    p_o, r_o, met_o = t_o.test(m_o, t_o.val_pos_list_df, t_o.val_interactions_t, K)
    assert list(met_o.columns) == list(met_r.columns)
    same = np.mean([len(set(a) & set(b)) / K for a, b in zip(met_o["top_rlvnt_itm"], met_r["top_rlvnt_itm"])])
    assert same > 0.999, same                                           # items swap only across fp32 near-ties
    assert abs(p_o - p_r) < 1e-4 and abs(r_o - r_r) < 2e-3
    # save_model / load (src/utils_v2.py:214-240): the reference's checkpoint code over the drop-in module
    tl_o.save_model(ck_o + "/LightGCN_best.pt", m_o, opt_o, p_o, r_o, epoch=0,
                    hyperparams={"latent_dim": DIM, "n_layers": LAYERS})
    best = torch.load(ck_o + "/LightGCN_best.pt", weights_only=False)
    assert list(best["model_state_dict"].keys()) == ["alpha", "embedding.weight"]
    assert tuple(best["model_state_dict"]["embedding.weight"].shape) == (n, DIM)


def _handler_module(lightgcn_module):
    """torchserve/lightgcn_handler.py imported unmodified; `ts` (TorchServe) is not installed: its
    BaseHandler base class is an empty stand-in."""
    root = os.path.dirname(reference_shim.REFERENCE_SRC)
    path = os.path.join(root, "torchserve", "lightgcn_handler.py")
    if not os.path.isfile(path):
        pytest.skip("handler source not staged")
    ts = types.ModuleType("ts"); th = types.ModuleType("ts.torch_handler"); bh = types.ModuleType("ts.torch_handler.base_handler")
    bh.BaseHandler = type("BaseHandler", (), {})
    saved = {k: sys.modules.get(k) for k in ("ts", "ts.torch_handler", "ts.torch_handler.base_handler", "lightgcn")}
    sys.modules.update({"ts": ts, "ts.torch_handler": th, "ts.torch_handler.base_handler": bh,
                        "lightgcn": lightgcn_module})
    try:
        spec = importlib.util.spec_from_file_location("ref_lightgcn_handler", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class _Context:
    def __init__(self, model_dir):
        self.manifest = {"model": {"serializedFile": "LightGCN_best.pt"}}
        self.system_properties = {"model_dir": model_dir, "gpu_id": 0}


def test_reference_torchserve_handler_over_the_dropin_module(tmp_path):
    import importlib.util  # noqa: F401
    from gnn_ecommerce_b200 import LightGCNService
    (tl_o, t_o, ck_o), (tl_r, t_r, ck_r) = _trainers(str(tmp_path))
    n = t_o.n_users + t_o.n_items
    torch.manual_seed(12)
    ref_lightgcn, _ = reference_shim.load_reference()
    m_r = ref_lightgcn.LightGCN(n, DIM, LAYERS)
    for ck in (ck_o, ck_r):                     # the model file + processed_train.csv the handler loads (:27-41)
        tl_r.save_model(ck + "LightGCN_best.pt", m_r, torch.optim.Adam(m_r.parameters(), 0.005), 0.0, 0.0, epoch=0,
                        hyperparams={"latent_dim": DIM, "n_layers": LAYERS})
    h_o = _handler_module(_dropin_module()).LightGCNHandler()
    h_r = _handler_module(ref_lightgcn).LightGCNHandler()
    h_o.initialize(_Context(ck_o)); h_r.initialize(_Context(ck_r))
    assert type(h_o.model) is _dropin_module().LightGCN
    request = [{"data": [0, 7, 19, 333, t_o.n_users - 1]}]
    out_o = h_o.postprocess(h_o.inference(h_o.preprocess(request)))
    out_r = h_r.postprocess(h_r.inference(h_r.preprocess(request)))
    assert list(out_o[0].keys()) == ["items"] and len(out_o[0]["items"]) == 5 and len(out_o[0]["items"][0]) == K
    agree = np.mean([len(set(a) & set(b)) / K for a, b in zip(out_o[0]["items"], out_r[0]["items"])])
    assert agree > 0.99, agree
    # the serving adapter of this repo (cached embeddings, CSR seen-lists): same contract, same lists
    frame = pd.read_csv(os.path.join(ck_o, "processed_train.csv"))
    svc = LightGCNService.from_train_frame(h_o.model, frame, DEV, k=K, handler_graph_quirk=True)
    assert torch.equal(svc.edge_index, h_o.edge_index)       # the handler builds its graph on un-offset item ids
    out_s = svc.handle(request)
    assert list(out_s[0].keys()) == ["items"]
    assert out_s[0]["items"] == out_o[0]["items"]
    with pytest.raises(IndexError):
        svc.inference([t_o.n_users])
    # default: the graph the model was trained on (what InferenceLightGCN scores, src/inference_lightgcn.py:18-45)
    svc2 = LightGCNService.from_train_frame(h_o.model, frame, DEV, k=K)
    assert torch.equal(svc2.edge_index, t_o.edge_index)
    users = request[0]["data"]
    mask = torch.index_select(h_r.i_m_matrix, 0, torch.as_tensor(users, device=h_r.i_m_matrix.device)).to_dense().cpu()
    want = h_r.model.recommendK(t_r.edge_index, t_r.edge_weight, t_r.n_users, t_r.n_items, mask, users, K)
    agree2 = np.mean([len(set(a) & set(b)) / K for a, b in zip(svc2.inference(users)["items"], want["top_rlvnt_itm"])])
    assert agree2 > 0.99, agree2


def test_device_sampler_from_the_reference_train_pos_list_frame(tmp_path):
    """`DeviceSampler.from_frame` on the `train_pos_list_df` the reference's own `prepare_val_test` builds
    (src/utils_v2.py:106-143): every triple obeys `batch_loader` (:168-181) -- distinct purchasers, the
    positive from the user's `item_id_idx_list`, the negative outside its `ignor_neg_list`."""
    from gnn_ecommerce_b200.sampler import DeviceSampler
    (_, t_o, _), _ = _trainers(str(tmp_path))
    frame = t_o.train_pos_list_df
    sampler = DeviceSampler.from_frame(frame, t_o.n_users, t_o.n_items, DEV, seed=3)
    pos_of = dict(zip(frame["user_id_idx"], frame["item_id_idx_list"]))
    ign_of = dict(zip(frame["user_id_idx"], frame["ignor_neg_list"]))
    seen_users = set()
    for _ in range(4):
        u, p, n = (x.cpu().numpy() for x in sampler.sample(256))
        assert len(set(u.tolist())) == 256 and set(u.tolist()) <= set(pos_of)
        for uu, pp, nn_ in zip(u, p, n):
            assert pp in pos_of[uu]
            assert nn_ not in set(ign_of[uu]) and t_o.n_users <= nn_ < t_o.n_users + t_o.n_items
        seen_users |= set(u.tolist())
    assert len(seen_users) > 600                    # batches differ
    with pytest.raises(ValueError):                 # random.sample: "Sample larger than population"
        sampler.sample(len(frame) + 1)
