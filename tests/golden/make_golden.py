"""Generate the committed golden vectors by running the REFERENCE ITSELF (its own models/*.py and
TGB/modules/*.py, unmodified, from /root/reference) through oracle/pyg_shim.py.

    python tests/golden/make_golden.py          # only works where /root/reference exists

Outputs (small, committed):
    tests/golden/loader_golden.pt   per-batch (n_id, edge_index, e_id) for K in {1,2} + final tables
    tests/golden/train_golden.pt    8 training steps of the reference CTAN / CTAN_tgb loop
                                    (train_link.py:237-282): logits, loss, final memory + params
    tests/golden/mrr_golden.pt      TGB evaluator MRR (numpy branch) on fixed scores
The reference's `nodes.sort()` is pinned to a stable sort while it runs (see pyg_shim.stable_sort).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyg_shim  # noqa: E402
import ctan_b200  # noqa: E402,F401
from ctan_b200 import synth  # noqa: E402


def golden_stream():
    st = synth.make_stream("wiki", n_events=1600, seed=0, n_nodes_scale=0.03, msg_dim=12)
    st["neg"] = synth.draw_negatives(st)[:, 0]
    return st


MODEL_KW = dict(edge_dim=12, memory_dim=16, time_dim=4, node_dim=1, epsilon=0.5, gamma=0.1, readout_hidden=8,
                mean_delta_t=1234.5, std_delta_t=4567.8, init_time=0)
CASES = [("ctan_k1", False, 1, None), ("ctan_k3_tanh", False, 3, "tanh"), ("tgb_k2", True, 2, None)]
B, S, STEPS = 200, 5, 8


def main():
    ns = pyg_shim.install()
    st = golden_stream()
    with pyg_shim.stable_sort():
        # ---------------- loader
        ld = ns.LastNeighborLoader(st["num_nodes"], 3)
        rec = {"S": 3, "B": 100, "batches": []}
        for lo in range(0, 1600, 100):
            s, d = st["src"][lo:lo + 100], st["dst"][lo:lo + 100]
            seeds = torch.cat([s, d, st["neg"][lo:lo + 100]]).unique()
            outs = {}
            for K in (1, 2):
                n_id = seeds
                for _ in range(K):
                    n_id, ei, eid = ld(n_id)
                outs[K] = (n_id.clone(), ei.clone(), eid.clone())
            rec["batches"].append(outs)
            ld.insert(s, d)
        rec["e_id"] = ld.e_id.clone()
        nb = ld.neighbors.clone()
        nb[ld.e_id < 0] = 0
        rec["neighbors"] = nb
        torch.save(rec, os.path.join(HERE, "loader_golden.pt"))

        # ---------------- training steps
        out = {}
        for name, tgb, K, final_act in CASES:
            torch.manual_seed(0)
            kw = dict(MODEL_KW, num_nodes=st["num_nodes"], num_iters=K, final_act=final_act)
            m = (ns.CTAN_tgb if tgb else ns.CTAN)(**kw)
            if tgb:
                m.layernorm = None                           # reference defect B3
            init = {k: v.clone() for k, v in m.state_dict().items() if not k.endswith("_assoc")}
            ld = ns.LastNeighborLoader(st["num_nodes"], S)
            helper = torch.zeros(st["num_nodes"], dtype=torch.long)
            opt = torch.optim.Adam(m.parameters(), lr=1e-3)
            crit = torch.nn.BCEWithLogitsLoss()
            steps = []
            for it in range(STEPS):
                lo, hi = it * B, (it + 1) * B
                s, d, t, msg, ng = st["src"][lo:hi], st["dst"][lo:hi], st["t"][lo:hi], st["msg"][lo:hi], st["neg"][lo:hi]
                opt.zero_grad()
                n_id = torch.cat([s, d, ng]).unique()
                for _ in range(m.num_gnn_layers):
                    n_id, ei, eid = ld(n_id)
                helper[n_id] = torch.arange(n_id.size(0))
                if tgb:
                    b = ns.TemporalData(src=torch.cat([s, s]), dst=torch.cat([d, ng]), t=t, msg=msg, x=st["x"], n_neg=B)
                else:
                    b = ns.TemporalData(src=s, dst=d, t=t, msg=msg, x=st["x"], neg_dst=ng)
                po, no, se, de = m(batch=b, n_id=n_id, msg=st["msg"][eid], t=st["t"][eid], edge_index=ei, id_mapper=helper)
                loss = crit(po, torch.ones_like(po)) + crit(no, torch.zeros_like(no))
                m.update(s, d, t, msg, se, de)
                ld.insert(s, d)
                loss.backward()
                grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None} if it == 0 else None
                opt.step()
                m.detach_memory()
                steps.append({"pos": po.detach().clone(), "neg": no.detach().clone(), "loss": loss.detach().clone(),
                              "grads": grads})
            out[name] = {"init": init, "steps": steps, "memory": m.memory.memory.clone(),
                         "last_update": m.memory.last_update.clone(),
                         "final": {k: v.detach().clone() for k, v in m.named_parameters()}}
        torch.save(out, os.path.join(HERE, "train_golden.pt"))

    # ---------------- MRR (TGB/tgb/linkproppred/evaluate.py:109-120, numpy branch, loaded standalone)
    import importlib.util, types
    src = open(os.path.join(pyg_shim.REFERENCE_ROOT, "TGB", "tgb", "linkproppred", "evaluate.py")).read()
    modu = types.ModuleType("_ref_evaluate")
    for name in ("tgb", "tgb.utils", "tgb.utils.info"):      # stub of the one symbol the file imports from tgb
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["tgb.utils.info"].DATA_EVAL_METRIC_DICT = {}
    # the file imports sklearn/torch/numpy only at module level besides tgb helpers it does not need here
    try:
        exec(compile(src, "evaluate.py", "exec"), modu.__dict__)
        ev = modu.Evaluator.__new__(modu.Evaluator)
        ev.k_value = 10
        g = torch.Generator().manual_seed(0)
        cases = []
        for n_neg in (1, 20, 999):
            pos = torch.randn(1, generator=g).numpy()
            neg = torch.randn(n_neg, generator=g).numpy()
            neg[: n_neg // 3] = pos[0]                       # ties
            r = ev._eval_hits_and_mrr(pos, neg.reshape(1, -1), type_info="numpy", k_value=10)
            cases.append({"pos": torch.from_numpy(pos.copy()), "neg": torch.from_numpy(neg.copy()), "mrr": float(r["mrr"])})
        torch.save(cases, os.path.join(HERE, "mrr_golden.pt"))
    except Exception as e:                                    # pragma: no cover
        print("could not exec the reference evaluator:", type(e).__name__, e)
    for f in ("loader_golden.pt", "train_golden.pt", "mrr_golden.pt"):
        p = os.path.join(HERE, f)
        if os.path.exists(p):
            print(f, os.path.getsize(p) // 1024, "KiB")


if __name__ == "__main__":
    main()
