"""C1 through `ctan_b200.train_sequence` (the reference's train / eval signatures, train_sequence.py:16, 68, on the
engine's sequence mode) against tests/golden/sequence_golden.pt -- what the reference's OWN train_sequence.train / eval
produced on the first 16 sequences of its shipped L=7 path-graph fixture -- and, when the reference files are present,
against the reference's own functions run live."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gold():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "sequence_golden.pt"), weights_only=False)


def _data(g, TD, dev):
    from oracle import ctan_oracle as O
    L, n = g["L"], g["n_seq"]
    src, dst, t = O.path_graph_events(L, n)
    x = g["x"].to(dev)
    return [TD(src=src[i * (L - 1):(i + 1) * (L - 1)].to(dev), dst=dst[i * (L - 1):(i + 1) * (L - 1)].to(dev),
               t=t[i * (L - 1):(i + 1) * (L - 1)].to(dev), msg=g["msg"][i].to(dev), y=g["y"][i:i + 1].to(dev), x=x)
            for i in range(n)]


def _model(g):
    from ctan_b200.models import CTAN, LastNeighborLoader
    mm = CTAN(**g["model_kw"]).to(DEV)
    missing = mm.load_state_dict({k: v for k, v in g["init"].items() if not k.endswith("_assoc")}, strict=False)
    assert not missing.unexpected_keys
    nn_ = g["model_kw"]["num_nodes"]
    return mm, LastNeighborLoader(nn_, 5, device=DEV), torch.zeros(nn_, dtype=torch.long, device=DEV)


@pytest.mark.parametrize("use_engine", [True, False])
def test_sequence_loops_match_the_reference_run(use_engine):
    from ctan_b200 import train_link as TL
    from ctan_b200 import train_sequence as TS
    g = _gold()
    mm, loader, helper = _model(g)
    data = _data(g, TL.TemporalData, DEV)
    if not use_engine:                                       # a sequence with its OWN node-feature table (differing only at a
        data[3].x = data[3].x.clone()                        # node it never touches): the engine declines, the module path runs
        data[3].x[0, -1, 0] += 1.0
    n, mb = g["n_seq"], g["meta_batch"]
    meta = [torch.arange(i, i + mb) for i in range(0, n, mb)]
    loaders = [TL.TemporalDataLoader(d, batch_size=1) for d in data]
    opt = torch.optim.Adam(mm.parameters(), lr=g["lr"], weight_decay=g["weight_decay"])
    crit = torch.nn.BCEWithLogitsLoss()
    assert TS.train(data, mm, opt, meta, loaders, crit, loader, helper, device=DEV) is None
    assert any(k[0] == "seq" for k in mm.__dict__.get("_ctan_engines", {})) == use_engine
    sd = mm.state_dict()
    for k, v in g["after_train"].items():
        if k.endswith("_assoc") or k not in sd or k.endswith("lin_key.bias"):   # (zero-gradient entry: Adam on noise)
            continue
        assert float((sd[k].cpu().to(v.dtype) - v).abs().max()) < 0.1 * g["lr"], k
    assert int(opt.state_dict()["state"][0]["step"]) == n // mb
    scores, (y_true, y_pred, y_conf) = TS.eval(data, mm, meta, loaders, crit, loader, helper, return_predictions=True,
                                               device=DEV)
    assert torch.equal(y_true, g["eval_true"])
    assert float((y_conf - g["eval_pred"]).abs().max()) < 2e-3
    assert abs(scores["loss"] - g["eval_loss"]) < 2e-3
    assert torch.equal(mm.memory.last_update.cpu(), g["last_update_after_eval"])
    assert float((mm.memory.memory.cpu() - g["memory_after_eval"]).abs().max()) < 2e-3


def test_shuffled_meta_batches_match_the_reference_live():
    """Processing order != storage order, ragged last meta-batch, two passes: against the reference's own functions."""
    from oracle import ref_loops
    if not ref_loops.available():
        pytest.skip("reference loop files not present")
    R = ref_loops.load()
    from oracle import pyg_shim
    from ctan_b200 import train_link as TL
    from ctan_b200 import train_sequence as TS
    g = _gold()
    n = g["n_seq"]
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(4))
    meta = [perm[0:6], perm[6:12], perm[12:16]]
    crit = torch.nn.BCEWithLogitsLoss()
    # reference, CPU
    rm = R.ns.CTAN(**g["model_kw"])
    rm.load_state_dict(g["init"])
    rdata = _data(g, R.TemporalData, "cpu")
    rl = R.ns.LastNeighborLoader(g["model_kw"]["num_nodes"], size=5)
    rh = torch.empty(g["model_kw"]["num_nodes"], dtype=torch.long)
    ropt = torch.optim.Adam(rm.parameters(), lr=g["lr"], weight_decay=g["weight_decay"])
    rloaders = [R.TemporalDataLoader(d, batch_size=1) for d in rdata]
    mm, ml, mh = _model(g)
    mdata = _data(g, TL.TemporalData, DEV)
    mopt = torch.optim.Adam(mm.parameters(), lr=g["lr"], weight_decay=g["weight_decay"])
    mloaders = [TL.TemporalDataLoader(d, batch_size=1) for d in mdata]
    with pyg_shim.stable_sort():
        for _ in range(2):
            R.train_sequence.train(rdata, rm, ropt, meta, rloaders, crit, rl, rh)
            TS.train(mdata, mm, mopt, meta, mloaders, crit, ml, mh, device=DEV)
        rs, (rt, _, rc) = R.train_sequence.eval(rdata, rm, meta, rloaders, crit, rl, rh, return_predictions=True)
    ms, (mt, _, mc) = TS.eval(mdata, mm, meta, mloaders, crit, ml, mh, return_predictions=True, device=DEV)
    assert torch.equal(mt, rt)
    assert float((mc - rc).abs().max()) < 5e-3
    assert abs(ms["loss"] - rs["loss"]) < 5e-3
    assert torch.equal(ml.e_id.cpu(), rl.e_id)
    assert torch.equal(mm.memory.last_update.cpu(), rm.memory.last_update)
    po = dict(rm.named_parameters())
    for k, p in mm.named_parameters():
        if "lin_skip" in k or k.endswith("lin_key.bias"):
            continue
        assert float((p.detach().cpu() - po[k].detach()).abs().max()) < 0.2 * g["lr"], k
