"""The reference's loop functions (train_link.py: tgb_train :206, tgb_test :24, train :287, eval :349).

Two directions:
  * the reference's OWN functions, imported unmodified (oracle/ref_loops.py: from /root/reference or its git-ignored
    mirror baseline/_ref/), driven on `ctan_b200.models` -- the two-line drop-in of INTEGRATION.md;
  * `ctan_b200.train_link` (same signatures, engines underneath) against the reference's own functions run on the
    reference's own model on the CPU, same seeds, same negatives.
Tolerances: losses / scores 5e-4 relative over <= 10 optimizer steps (1e-4 per step, fp32 path), integer state exact."""
import contextlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
B, S = 50, 5


def _ref():
    from oracle import ref_loops
    if not ref_loops.available():
        pytest.skip("reference loop files not present (neither /root/reference nor baseline/_ref/)")
    return ref_loops, ref_loops.load()


@pytest.fixture(autouse=True)
def _stable_reference_sort():
    """The reference's LastNeighborLoader.insert sorts without stable=True: pin it to the stable behaviour (the
    contract) while the reference's own loops run on the CPU."""
    from oracle import pyg_shim
    with pyg_shim.stable_sort():
        yield


@contextlib.contextmanager
def fixed_randint(neg):
    """Make `torch.randint(lo, hi, (n,), ...)` hand out consecutive slices of `neg` (on the requested device) while
    a reference loop runs: CPU and CUDA generators differ, the loops must see the same negatives."""
    orig, pos = torch.randint, [0]

    def fake(*a, **k):
        n = (a[2] if len(a) > 2 else k["size"])[0]
        out = neg[pos[0]:pos[0] + n].to(k.get("device", "cpu"))
        pos[0] += n
        return out
    torch.randint = fake
    try:
        yield
    finally:
        torch.randint = orig


def _data(R_or_ours, st, n, dev, y=None):
    d = R_or_ours.TemporalData(src=st["src"][:n].to(dev), dst=st["dst"][:n].to(dev), t=st["t"][:n].to(dev),
                               msg=st["msg"][:n].to(dev))
    d.x = st["x"].to(dev)
    if y is not None:
        d.y = y[:n].to(dev)
    return d


def _kw(st, tgb=False, K=1, out=1, final_act=None):
    kw = dict(num_nodes=st["num_nodes"], edge_dim=st["msg"].size(1), memory_dim=32, time_dim=8, node_dim=1,
              num_iters=K, epsilon=0.5, gamma=0.1, mean_delta_t=np.float64(1234.5), std_delta_t=np.float64(4567.8),
              final_act=final_act)
    if not tgb:
        kw.update(readout_hidden=16, readout_out=out)
    return kw


def _ref_model(R, st, tgb, **kw):
    torch.manual_seed(0)
    m = (R.ns.CTAN_tgb if tgb else R.ns.CTAN)(**_kw(st, tgb, **kw))
    m.layernorm = None                                       # reference defect B3 (attribute read, never defined)
    return m, R.ns.LastNeighborLoader(st["num_nodes"], size=S), torch.empty(st["num_nodes"], dtype=torch.long)


def _our_model(st, tgb, **kw):
    from ctan_b200.models import CTAN, CTAN_tgb, LastNeighborLoader
    torch.manual_seed(0)
    m = (CTAN_tgb if tgb else CTAN)(**_kw(st, tgb, **kw)).to(DEV)
    return m, LastNeighborLoader(st["num_nodes"], S, device=DEV), torch.empty(st["num_nodes"], dtype=torch.long, device=DEV)


def _close(a, b, tol=5e-4):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.all(np.abs(a - b) <= tol * np.abs(b) + 1e-6), (a, b)


def _state_equal(mm, ld, rm, rl):
    assert torch.equal(ld.e_id.cpu(), rl.e_id)
    assert torch.equal(mm.memory.last_update.cpu(), rm.memory.last_update)
    mo = rm.memory.memory
    assert float((mm.memory.memory.cpu() - mo).abs().max()) <= 1e-3 * float(mo.abs().max()) + 1e-6


def test_reference_own_tgb_train_runs_on_our_models(small_stream):
    """The reference's own tgb_train (unmodified) with ctan_b200.models swapped in == the same call on the
    reference's own classes."""
    RL, R = _ref()
    st, n = small_stream, 8 * B + 30
    crit = torch.nn.BCEWithLogitsLoss()
    rm, rl, rh = _ref_model(R, st, True, K=2, final_act="tanh")
    ropt = torch.optim.Adam(rm.parameters(), lr=1e-3)
    rd = _data(R, st, n, "cpu")
    with fixed_randint(st["neg"]):
        ref = R.train_link.tgb_train(rd, rm, ropt, R.TemporalDataLoader(rd, batch_size=B), crit, rl, rh, device="cpu")
    mm, ml, mh = _our_model(st, True, K=2, final_act="tanh")
    mopt = torch.optim.Adam(mm.parameters(), lr=1e-3)
    md = _data(R, st, n, DEV)
    with fixed_randint(st["neg"]):
        mine = R.train_link.tgb_train(md, mm, mopt, R.TemporalDataLoader(md, batch_size=B), crit, ml, mh, device=DEV)
    assert len(mine) == len(ref) == 9
    _close(mine, ref)
    _state_equal(mm, ml, rm, rl)


@pytest.mark.parametrize("tgb", [True, False])
def test_engine_loops_match_reference_loops(small_stream, tgb):
    """ctan_b200.train_link.tgb_train / train (engine + short tail batch) == the reference's own function, two epochs
    with one optimizer (Adam's step count carries across the per-epoch reset), then eval / tgb_test on the next split."""
    from ctan_b200 import train_link as TL
    RL, R = _ref()
    st, n_tr, n_ev = small_stream, 8 * B + 30, 4 * B + 20
    n = n_tr + n_ev
    crit = torch.nn.BCEWithLogitsLoss()
    rm, rl, rh = _ref_model(R, st, tgb)
    ropt = torch.optim.Adam(rm.parameters(), lr=2e-3)
    rd = _data(R, st, n, "cpu")
    r_tr, r_ev = rd[:n_tr], rd[n_tr:n]
    mm, ml, mh = _our_model(st, tgb)
    mopt = torch.optim.Adam(mm.parameters(), lr=2e-3)
    md = _data(TL, st, n, DEV)
    m_tr, m_ev = md[:n_tr], md[n_tr:n]
    for epoch in range(2):
        if tgb:
            with fixed_randint(st["neg"]):
                ref = R.train_link.tgb_train(rd, rm, ropt, R.TemporalDataLoader(r_tr, batch_size=B), crit, rl, rh)
            with fixed_randint(st["neg"]):
                mine = TL.tgb_train(md, mm, mopt, TL.TemporalDataLoader(m_tr, batch_size=B), crit, ml, mh, device=DEV)
        else:
            ref = R.train_link.train(rd, rm, ropt, R.TemporalDataLoader(r_tr, batch_size=B), crit, rl, rh,
                                     train_neg_sampler=RL.FixedNegatives(st["neg"]))
            mine = TL.train(md, mm, mopt, TL.TemporalDataLoader(m_tr, batch_size=B), crit, ml, mh,
                            train_neg_sampler=RL.FixedNegatives(st["neg"]), device=DEV)
        assert len(mine) == len(ref) == 9
        _close(mine, ref, 1e-3 if epoch else 5e-4)
        _state_equal(mm, ml, rm, rl)
    # the optimizer the caller holds IS the engine's Adam state (checkpointable the reference's way)
    sd = mopt.state_dict()
    assert int(sd["state"][0]["step"]) == 18 and sd["state"][0]["exp_avg"].abs().sum() > 0
    rsd = ropt.state_dict()["state"]
    names = [k for k, _ in mm.named_parameters()]
    rnames = [k for k, _ in rm.named_parameters()]
    for i, k in enumerate(names):
        if "lin_skip" in k:
            continue
        j = rnames.index(k)
        a, b = sd["state"][i]["exp_avg_sq"].cpu(), rsd[j]["exp_avg_sq"]
        assert float((a - b).abs().max()) <= 2e-3 * float(b.abs().max()) + 1e-12, k
    # evaluation continues from the training state (no reset), on the next split
    if tgb:
        negs = torch.randint(st["min_dst"], st["max_dst"] + 1, (n, 7), generator=torch.Generator().manual_seed(3))
        ev = R.Evaluator(name="tgbl-wiki")
        ref_s, _ = R.train_link.tgb_test(rd, rm, R.TemporalDataLoader(r_ev, batch_size=B),
                                         RL.FixedEvalNegatives(negs, start=n_tr), rl, "val", rh, ev, "mrr")
        mine_s, _ = TL.tgb_test(md, mm, TL.TemporalDataLoader(m_ev, batch_size=B), RL.FixedEvalNegatives(negs, start=n_tr),
                                ml, "val", mh, ev, "mrr", device=DEV)
        assert abs(mine_s["mrr"] - ref_s["mrr"]) <= 5e-3, (mine_s, ref_s)
        _state_equal(mm, ml, rm, rl)
    else:
        ref_s, ref_p = R.train_link.eval(rd, rm, R.TemporalDataLoader(r_ev, batch_size=B), crit, rl, rh,
                                         neg_sampler=RL.FixedNegatives(st["neg"], start=n_tr), return_predictions=True)
        mine_s, mine_p = TL.eval(md, mm, TL.TemporalDataLoader(m_ev, batch_size=B), crit, ml, mh,
                                 neg_sampler=RL.FixedNegatives(st["neg"], start=n_tr), return_predictions=True, device=DEV)
        assert torch.equal(mine_p[0], ref_p[0])
        assert float((mine_p[2] - ref_p[2]).abs().max()) <= 2e-3 * float(ref_p[2].abs().max())
        for k in ("auc", "ap", "loss"):
            assert abs(float(mine_s[k]) - float(ref_s[k])) <= 5e-3, (k, mine_s[k], ref_s[k])
        _state_equal(mm, ml, rm, rl)


def test_tgb_test_train_split_matches_reference(small_stream):
    from ctan_b200 import train_link as TL
    RL, R = _ref()
    st, n = small_stream, 5 * B + 10
    rm, rl, rh = _ref_model(R, st, True)
    mm, ml, mh = _our_model(st, True)
    rd, md = _data(R, st, n, "cpu"), _data(TL, st, n, DEV)
    ev = R.Evaluator(name="tgbl-wiki")
    rm.reset_memory(); rl.reset_state(); mm.reset_memory(); ml.reset_state()
    with fixed_randint(st["neg"]):
        ref_s, _ = R.train_link.tgb_test(rd, rm, R.TemporalDataLoader(rd, batch_size=B), None, rl, "train", rh, ev, "mrr")
    with fixed_randint(st["neg"]):
        mine_s, _ = TL.tgb_test(md, mm, TL.TemporalDataLoader(md, batch_size=B), None, ml, "train", mh, ev, "mrr", device=DEV)
    assert abs(mine_s["mrr"] - ref_s["mrr"]) <= 2e-3, (mine_s, ref_s)
    _state_equal(mm, ml, rm, rl)


def test_multiclass_train_matches_reference(small_stream):
    """C4's head: 21 classes, CrossEntropyLoss, no negatives (train_link.py:304-306, 329, 708-709) through the
    engine's labelled path."""
    from ctan_b200 import train_link as TL
    RL, R = _ref()
    st, n = small_stream, 8 * B + 30
    y = torch.randint(0, 21, (st["src"].numel(),), generator=torch.Generator().manual_seed(5))
    crit = torch.nn.CrossEntropyLoss()
    rm, rl, rh = _ref_model(R, st, False, out=21, K=2)
    ropt = torch.optim.Adam(rm.parameters(), lr=2e-3)
    rd = _data(R, st, n, "cpu", y)
    ref = R.train_link.train(rd, rm, ropt, R.TemporalDataLoader(rd, batch_size=B), crit, rl, rh)
    mm, ml, mh = _our_model(st, False, out=21, K=2)
    mopt = torch.optim.Adam(mm.parameters(), lr=2e-3)
    md = _data(TL, st, n, DEV, y)
    mine = TL.train(md, mm, mopt, TL.TemporalDataLoader(md, batch_size=B), crit, ml, mh, device=DEV)
    assert any(k[0] == "step" for k in mm._ctan_engines), "the engine path did not run"
    _close(mine, ref)
    _state_equal(mm, ml, rm, rl)
    po = dict(rm.named_parameters())
    for k, p in mm.named_parameters():
        if "lin_skip" not in k:
            assert float((p.detach().cpu() - po[k].detach()).abs().max()) <= 5e-4 * float(po[k].abs().max()) + 1e-6, k


@pytest.mark.parametrize("cname", ["MSELoss", "L1Loss"])
def test_regression_train_matches_reference(small_stream, cname):
    from ctan_b200 import train_link as TL
    RL, R = _ref()
    st, n = small_stream, 6 * B
    y = torch.randn(st["src"].numel(), 1, generator=torch.Generator().manual_seed(6))
    crit = getattr(torch.nn, cname)()
    rm, rl, rh = _ref_model(R, st, False)
    ropt = torch.optim.Adam(rm.parameters(), lr=1e-3)
    rd = _data(R, st, n, "cpu", y)
    ref = R.train_link.train(rd, rm, ropt, R.TemporalDataLoader(rd, batch_size=B), crit, rl, rh)
    mm, ml, mh = _our_model(st, False)
    mopt = torch.optim.Adam(mm.parameters(), lr=1e-3)
    md = _data(TL, st, n, DEV, y)
    mine = TL.train(md, mm, mopt, TL.TemporalDataLoader(md, batch_size=B), crit, ml, mh, device=DEV)
    _close(mine, ref)
    _state_equal(mm, ml, rm, rl)


def test_checkpoint_resume_through_the_optimizer(small_stream):
    """train_link.py:587-603 / 497-511: save model + optimizer state dicts after an engine epoch, resume in fresh
    objects, train another epoch == two uninterrupted epochs."""
    from ctan_b200 import train_link as TL
    st, n = small_stream, 6 * B
    crit = torch.nn.BCEWithLogitsLoss()

    def fresh():
        m, l, h = _our_model(st, True)
        return m, l, h, torch.optim.Adam(m.parameters(), lr=2e-3)
    md = _data(TL, st, n, DEV)
    m1, l1, h1, o1 = fresh()
    for _ in range(2):
        with fixed_randint(st["neg"]):
            full = TL.tgb_train(md, m1, o1, TL.TemporalDataLoader(md, batch_size=B), crit, l1, h1, device=DEV)
    m2, l2, h2, o2 = fresh()
    with fixed_randint(st["neg"]):
        TL.tgb_train(md, m2, o2, TL.TemporalDataLoader(md, batch_size=B), crit, l2, h2, device=DEV)
    ck = {"model": {k: v.clone() for k, v in m2.state_dict().items()}, "opt": o2.state_dict()}
    ck = torch.load(__import__("io").BytesIO(_dump(ck)), map_location=DEV, weights_only=False)
    m3, l3, h3, o3 = fresh()
    m3.load_state_dict(ck["model"])
    o3.load_state_dict(ck["opt"])
    with fixed_randint(st["neg"]):
        resumed = TL.tgb_train(md, m3, o3, TL.TemporalDataLoader(md, batch_size=B), crit, l3, h3, device=DEV)
    _close(resumed, full, 1e-5)
    for (k, a), (_, b) in zip(m3.named_parameters(), m1.named_parameters()):
        if k.endswith("lin_key.bias"):
            continue        # its true gradient is exactly 0 (softmax is shift invariant): Adam normalises pure atomic-order noise
        assert float((a - b).abs().max()) <= 1e-6 + 1e-4 * float(b.abs().max()), k


def _dump(obj):
    import io
    buf = io.BytesIO()
    torch.save(obj, buf)
    return buf.getvalue()


def test_compute_stats_reference_signature(small_stream):
    """utils.compute_stats(data, split, init_time, sequence_prediction) (the caller at main.py:302 /
    main_tgb_ctan.py:270), imported unmodified, against ctan_b200.stats.compute_stats_from_data on the same objects."""
    from ctan_b200.stats import compute_stats_from_data
    RL, R = _ref()
    st, n = small_stream, 1500
    d = R.TemporalData(src=st["src"][:n], dst=st["dst"][:n], t=st["t"][:n], msg=st["msg"][:n])
    ref = R.utils.compute_stats(d, (0.15, 0.15), 0)
    mine = compute_stats_from_data(d, (0.15, 0.15), 0, device=DEV)
    assert abs(mine[0] - float(ref[0])) <= 1e-9 * abs(float(ref[0])) and abs(mine[1] - float(ref[1])) <= 1e-9 * float(ref[1])

    class Seqs:                                              # the sequence case: a dataset object whose split returns lists
        def __init__(self):
            self.data = [R.TemporalData(src=torch.arange(i * 7, i * 7 + 6), dst=torch.arange(i * 7 + 1, i * 7 + 7),
                                        t=torch.arange(6)) for i in range(20)]

        def train_val_test_split(self, val_ratio, test_ratio):
            a = int(len(self.data) * (1 - val_ratio - test_ratio))
            return self.data[:a], self.data[a:a + 3], self.data[a + 3:]
    ref = R.utils.compute_stats(Seqs(), (0.15, 0.15), 0, sequence_prediction=True)
    mine = compute_stats_from_data(Seqs(), (0.15, 0.15), 0, sequence_prediction=True, device=DEV)
    assert abs(mine[0] - float(ref[0])) <= 1e-12 + 1e-9 * abs(float(ref[0])) and abs(mine[1] - float(ref[1])) <= 1e-9 * float(ref[1])
