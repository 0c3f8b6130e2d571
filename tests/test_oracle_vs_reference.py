"""Live cross-check of the oracle against the reference's own code (run through oracle/pyg_shim.py).
Only runs where /root/reference exists (the build container); skipped on the GPU box."""
import numpy as np
import pytest
import torch

from oracle import ctan_oracle as O
from oracle import pyg_shim

pytestmark = pytest.mark.skipif(not pyg_shim.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return pyg_shim.install()


@pytest.mark.parametrize("S", [1, 3, 5, 32])
def test_loader_replay(ref, small_stream, S):
    st = small_stream
    with pyg_shim.stable_sort():
        r = ref.LastNeighborLoader(st["num_nodes"], S)
        o = O.TorchNeighborLoader(st["num_nodes"], S)
        sp = O.NeighborTableSpec(st["num_nodes"], S)
        for lo in range(0, 2400, 200):
            s, d = st["src"][lo:lo + 200], st["dst"][lo:lo + 200]
            seeds = torch.cat([s, d, st["neg"][lo:lo + 200]]).unique()
            for K in (1, 2, 3):
                for a, b in zip(O.khop_lookup(r, seeds, K), O.khop_lookup(o, seeds, K)):
                    assert torch.equal(a, b)
            for a, b in zip(r(seeds), sp.lookup(seeds.numpy())):
                assert np.array_equal(a.numpy(), b)
            r.insert(s, d); o.insert(s, d); sp.insert(s.numpy(), d.numpy())
            assert torch.equal(r.e_id, o.e_id) and np.array_equal(r.e_id.numpy(), sp.e_id)
            m = r.e_id >= 0
            assert torch.equal(r.neighbors[m], o.neighbors[m])
            assert np.array_equal(r.neighbors[m].numpy(), sp.neighbors[m.numpy()])


def test_unstable_sort_is_the_only_divergence(ref, small_stream):
    """Documents WHY the stable-sort pin exists: without it the reference's own overflow result
    depends on torch's unstable CPU sort; with <= S events per node and batch it never matters."""
    st = small_stream
    r = ref.LastNeighborLoader(st["num_nodes"], 400)       # S >= 2B: overflow impossible
    o = O.TorchNeighborLoader(st["num_nodes"], 400)
    for lo in range(0, 1000, 200):
        s, d = st["src"][lo:lo + 200], st["dst"][lo:lo + 200]
        r.insert(s, d); o.insert(s, d)
    assert torch.equal(r.e_id, o.e_id)


@pytest.mark.parametrize("tgb,K,final_act", [(False, 1, None), (False, 3, "tanh"), (True, 2, "tanh")])
def test_model_train_steps(ref, small_stream, tgb, K, final_act):
    st = small_stream
    kw = dict(num_nodes=st["num_nodes"], edge_dim=172, memory_dim=32, time_dim=8, node_dim=1, num_iters=K,
              epsilon=0.5, gamma=0.1, readout_hidden=16, mean_delta_t=np.float64(1234.5),
              std_delta_t=np.float64(4567.8), init_time=torch.tensor(0), final_act=final_act)
    torch.manual_seed(0)
    rm = (ref.CTAN_tgb if tgb else ref.CTAN)(**kw)
    if tgb:
        rm.layernorm = None
    torch.manual_seed(0)
    om = O.OracleCTAN(**kw, tgb=tgb)
    sr, so = rm.state_dict(), om.state_dict()
    assert set(sr) == set(so)
    for k in sr:
        if not k.endswith("_assoc"):
            assert torch.equal(sr[k], so[k]), k
    with pyg_shim.stable_sort():
        lr_, lo_ = ref.LastNeighborLoader(st["num_nodes"], 5), O.TorchNeighborLoader(st["num_nodes"], 5)
        hr = torch.zeros(st["num_nodes"], dtype=torch.long); ho = hr.clone()
        opr, opo = torch.optim.Adam(rm.parameters(), lr=1e-3), torch.optim.Adam(om.parameters(), lr=1e-3)
        crit = torch.nn.BCEWithLogitsLoss()
        for lo in range(0, 2000, 200):
            hi = lo + 200
            s, d, t, msg, ng = st["src"][lo:hi], st["dst"][lo:hi], st["t"][lo:hi], st["msg"][lo:hi], st["neg"][lo:hi]
            opr.zero_grad()
            n_id = torch.cat([s, d, ng]).unique()
            for _ in range(rm.num_gnn_layers):
                n_id, ei, eid = lr_(n_id)
            hr[n_id] = torch.arange(n_id.size(0))
            if tgb:
                b = ref.TemporalData(src=torch.cat([s, s]), dst=torch.cat([d, ng]), t=t, msg=msg, x=st["x"], n_neg=200)
            else:
                b = ref.TemporalData(src=s, dst=d, t=t, msg=msg, x=st["x"], neg_dst=ng)
            po, no, se, de = rm(batch=b, n_id=n_id, msg=st["msg"][eid], t=st["t"][eid], edge_index=ei, id_mapper=hr)
            loss = crit(po, torch.ones_like(po)) + crit(no, torch.zeros_like(no))
            rm.update(s, d, t, msg, se, de); lr_.insert(s, d); loss.backward(); opr.step(); rm.detach_memory()
            l2, po2, no2 = O.train_step_ref(om, lo_, ho, opo, st, lo, hi, ng)
            assert torch.allclose(loss.detach(), l2, rtol=1e-6, atol=1e-7)
            assert torch.allclose(po.detach(), po2, rtol=1e-5, atol=1e-6)
        assert torch.equal(rm.memory.last_update, om.memory.last_update)
        assert torch.allclose(rm.memory.memory, om.memory.memory, rtol=1e-5, atol=1e-6)


def test_same_seed_init_of_cuda_modules_matches_reference(ref):
    """The product's module mirror draws its parameters exactly like the reference (App. A.7)."""
    import ctan_b200  # noqa: F401
    from ctan_b200.models import CTAN, CTAN_tgb
    kw = dict(num_nodes=64, edge_dim=5, memory_dim=12, time_dim=3, node_dim=2, num_iters=2, readout_hidden=6)
    for cls_r, cls_m in ((ref.CTAN, CTAN), (ref.CTAN_tgb, CTAN_tgb)):
        torch.manual_seed(7); a = cls_r(**kw)
        torch.manual_seed(7); b = cls_m(**kw)
        sa, sb = a.state_dict(), b.state_dict()
        assert set(sa) == set(sb)
        for k in sa:
            if not k.endswith("_assoc"):
                assert torch.equal(sa[k], sb[k]), k


@pytest.mark.parametrize("predict_dst", [False, True])
def test_multiclass_and_sequence_heads(ref, predict_dst):
    """C4 (Pascal-VOC: no negatives, 21 classes, CrossEntropy, train_link.py:297-343) and C1's SequencePredictor head:
    same-seed init, logits, loss and every gradient of the reference model against the oracle."""
    g = torch.Generator().manual_seed(4)
    num_nodes, n, B, D, Fn, out = 300, 512, 128, 37, 14, (1 if predict_dst else 21)
    src = torch.randint(0, num_nodes, (n,), generator=g)
    dst = torch.randint(0, num_nodes, (n,), generator=g)
    t = torch.sort(torch.randint(0, 5000, (n,), generator=g)).values
    msg = torch.randn(n, 1, generator=g)
    x = torch.randn(num_nodes, Fn, generator=g)
    y = torch.randint(0, 21, (n,), generator=g)
    kw = dict(num_nodes=num_nodes, edge_dim=1, memory_dim=D, time_dim=1, node_dim=Fn, num_iters=3, epsilon=0.1,
              gamma=0.1, readout_hidden=D // 2, readout_out=out, mean_delta_t=30.0, std_delta_t=50.0,
              predict_dst=predict_dst)
    torch.manual_seed(1)
    rm = ref.CTAN(**kw)
    torch.manual_seed(1)
    om = O.OracleCTAN(**kw)
    for (ka, a), (kb, b) in zip(rm.state_dict().items(), om.state_dict().items()):
        assert ka == kb and (ka.endswith("_assoc") or torch.equal(a, b)), ka
    with pyg_shim.stable_sort():
        lr_, lo_ = ref.LastNeighborLoader(num_nodes, 5), O.TorchNeighborLoader(num_nodes, 5)
        hr, ho = torch.zeros(num_nodes, dtype=torch.long), torch.zeros(num_nodes, dtype=torch.long)
        for it in range(3):
            a, b = it * B, (it + 1) * B
            res = []
            for model, loader, helper in ((rm, lr_, hr), (om, lo_, ho)):
                model.zero_grad()
                n_id = torch.cat([src[a:b], dst[a:b]]).unique()
                for _ in range(model.num_gnn_layers):
                    n_id, ei, eid = loader(n_id)
                helper[n_id] = torch.arange(n_id.numel())
                pos, neg, se, de = model(O.Batch(src=src[a:b], dst=dst[a:b], x=x), n_id, msg[eid], t[eid], ei, helper)
                assert neg is None
                loss = (torch.nn.BCEWithLogitsLoss()(pos.view(-1), (y[a:b] % 2).float()) if predict_dst
                        else torch.nn.CrossEntropyLoss()(pos, y[a:b]))
                model.update(src[a:b], dst[a:b], t[a:b], msg[a:b], se, de)
                loader.insert(src[a:b], dst[a:b])
                loss.backward()
                res.append((pos.detach(), loss.detach(), {k: p.grad.clone() for k, p in model.named_parameters()
                                                          if p.grad is not None}))
            (p_r, l_r, g_r), (p_o, l_o, g_o) = res
            assert torch.allclose(p_r, p_o, rtol=1e-5, atol=1e-6) and torch.allclose(l_r, l_o, rtol=1e-6)
            assert set(g_r) == set(g_o)
            for k in g_r:
                assert torch.allclose(g_r[k], g_o[k], rtol=1e-4, atol=1e-6), k
