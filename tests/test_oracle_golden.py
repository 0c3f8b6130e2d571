"""Pins the oracle (oracle/ctan_oracle.py) to golden vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import ctan_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


def _stream():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, mod.golden_stream()


@pytest.mark.parametrize("impl", ["torch", "spec"])
def test_loader_matches_reference_golden(impl):
    mod, st = _stream()
    rec = torch.load(os.path.join(G, "loader_golden.pt"))
    S, B = rec["S"], rec["B"]
    ld = O.TorchNeighborLoader(st["num_nodes"], S) if impl == "torch" else O.NeighborTableSpec(st["num_nodes"], S)
    for bi, lo in enumerate(range(0, 1600, B)):
        s, d = st["src"][lo:lo + B], st["dst"][lo:lo + B]
        seeds = torch.cat([s, d, st["neg"][lo:lo + B]]).unique()
        for K in (1, 2):
            if impl == "torch":
                got = O.khop_lookup(ld, seeds, K)
            else:
                n_id = seeds.numpy()
                for _ in range(K):
                    n_id, ei, eid = ld.lookup(n_id)
                got = (torch.from_numpy(n_id), torch.from_numpy(ei), torch.from_numpy(eid))
            for a, b in zip(got, rec["batches"][bi][K]):
                assert torch.equal(a, b), (bi, K)
        if impl == "torch":
            ld.insert(s, d)
        else:
            ld.insert(s.numpy(), d.numpy())
    e = ld.e_id if impl == "torch" else torch.from_numpy(ld.e_id)
    nb = ld.neighbors if impl == "torch" else torch.from_numpy(ld.neighbors)
    assert torch.equal(e, rec["e_id"])
    assert torch.equal(nb[e >= 0], rec["neighbors"][e >= 0])


@pytest.mark.parametrize("case", ["ctan_k1", "ctan_k3_tanh", "tgb_k2"])
def test_training_matches_reference_golden(case):
    mod, st = _stream()
    gold = torch.load(os.path.join(G, "train_golden.pt"))[case]
    name, tgb, K, final_act = [c for c in mod.CASES if c[0] == case][0]
    torch.manual_seed(0)
    m = O.OracleCTAN(**dict(mod.MODEL_KW, num_nodes=st["num_nodes"], num_iters=K, final_act=final_act), tgb=tgb)
    for k, v in gold["init"].items():                       # same-seed init equals the reference's
        assert torch.equal(m.state_dict()[k], v), k
    ld = O.TorchNeighborLoader(st["num_nodes"], mod.S)
    helper = torch.zeros(st["num_nodes"], dtype=torch.long)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    for it, ref in enumerate(gold["steps"]):
        lo, hi = it * mod.B, (it + 1) * mod.B
        loss, po, no = O.train_step_ref(m, ld, helper, opt, st, lo, hi, st["neg"][lo:hi])
        if it == 0:
            grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
            assert set(grads) == set(ref["grads"])
            for k in grads:
                assert torch.allclose(grads[k], ref["grads"][k], rtol=1e-5, atol=1e-7), k
        assert torch.allclose(po, ref["pos"], rtol=1e-5, atol=1e-6), it
        assert torch.allclose(no, ref["neg"], rtol=1e-5, atol=1e-6), it
        assert torch.allclose(loss, ref["loss"], rtol=1e-6, atol=1e-7), it
    assert torch.equal(m.memory.last_update, gold["last_update"])
    assert torch.allclose(m.memory.memory, gold["memory"], rtol=1e-5, atol=1e-6)
    for k, p in m.named_parameters():
        assert torch.allclose(p, gold["final"][k], rtol=1e-5, atol=1e-6), k


def test_mrr_matches_reference_evaluator():
    for c in torch.load(os.path.join(G, "mrr_golden.pt")):
        assert abs(O.mrr_ref(c["pos"].numpy(), c["neg"].numpy()) - c["mrr"]) < 1e-7


def test_memory_update_tie_rule():
    """First position wins on equal times (torch-scatter CPU arg-max), last_update = max t."""
    mem, lu = torch.zeros(6, 2), torch.zeros(6, dtype=torch.long)
    src, dst = torch.tensor([1, 1, 2]), torch.tensor([3, 1, 1])
    t = torch.tensor([5, 5, 4])
    se = torch.tensor([[1., 1], [2, 2], [3, 3]])
    de = torch.tensor([[4., 4], [5, 5], [6, 6]])
    O.memory_update_ref(mem, lu, src, dst, t, se, de)
    assert lu.tolist() == [0, 5, 4, 5, 0, 0]
    assert mem[1].tolist() == [1, 1] and mem[2].tolist() == [3, 3] and mem[3].tolist() == [4, 4]


# ------------------------------------------------------------------ C1: the reference's shipped fixture
def _seq_golden():
    return torch.load(os.path.join(G, "sequence_golden.pt"), weights_only=False)


def test_delta_t_stats_known_answers():
    """`delta_t_stats.pkl` of the shipped label-sequence fixture = the reference's own known answers of
    utils.compute_stats for L in {3,...,20} (train split = first int(1000*(1-.15-.15)) sequences)."""
    g = _seq_golden()
    assert sorted(g["stats"]) == [3, 5, 7, 9, 11, 13, 15, 20]
    n_train = int(1000 * (1 - g["split"][0] - g["split"][1]))
    for L, (mean, std) in g["stats"].items():
        s, d, t = O.path_graph_events(L, n_train)
        m, sd = O.delta_t_stats_ref(s, d, t, init_time=0)
        assert abs(m - mean) < 1e-12 and abs(sd - std) < 1e-12, (L, m, sd)


def test_oracle_sequence_loop_matches_reference_run():
    """The reference's train_sequence.train + eval, run verbatim on the first 16 sequences of its shipped L=7
    dataset (tests/golden/make_sequence_golden.py): parameters after the optimizer steps, logits, memory."""
    g = _seq_golden()
    L, n = g["L"], g["n_seq"]
    om = O.OracleCTAN(**g["model_kw"])
    om.load_state_dict(g["init"])
    src, dst, t = O.path_graph_events(L, n)
    seqs = [dict(src=src[i * (L - 1):(i + 1) * (L - 1)], dst=dst[i * (L - 1):(i + 1) * (L - 1)],
                 t=t[i * (L - 1):(i + 1) * (L - 1)], msg=g["msg"][i], y=g["y"][i:i + 1]) for i in range(n)]
    meta = [range(i, i + g["meta_batch"]) for i in range(0, n, g["meta_batch"])]
    loader = O.TorchNeighborLoader(g["model_kw"]["num_nodes"], 5)
    helper = torch.zeros(g["model_kw"]["num_nodes"], dtype=torch.long)
    opt = torch.optim.Adam(om.parameters(), lr=g["lr"], weight_decay=g["weight_decay"])
    O.sequence_pass_ref(om, loader, helper, seqs, g["x"], meta, optimizer=opt)
    sd = om.state_dict()
    for k, v in g["after_train"].items():
        if not k.endswith("_assoc"):            # scratch table (torch.empty in the reference)
            assert torch.allclose(sd[k].to(v.dtype), v, rtol=1e-5, atol=1e-6), k
    pred, true = O.sequence_pass_ref(om, loader, helper, seqs, g["x"], meta)
    assert torch.equal(true, g["eval_true"])
    assert torch.allclose(pred, g["eval_pred"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(om.memory.memory, g["memory_after_eval"], rtol=1e-5, atol=1e-6)
    assert torch.equal(om.memory.last_update, g["last_update_after_eval"])
