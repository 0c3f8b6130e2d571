"""Golden vectors for config C1 (Temporal Path Graph sequence classification) from the reference's OWN
fixture and OWN code, generated where /root/reference exists:

    python tests/golden/make_sequence_golden.py

* the shipped fixture `data/label_sequence.tar.xz` holds, per sequence length L, the pre-generated dataset
  and the known answers of `utils.compute_stats` (`delta_t_stats.pkl`) -> `stats[L] = (mean, std)`;
* the first N_SEQ sequences of the L=7 dataset (real `msg`, `x`, `y`) are fed to the reference's
  `train_sequence.train` / `eval` (imported verbatim, PyG pieces from oracle/pyg_shim.py) with the reference
  `CTAN` (conf_sequence.py shapes): predictions, loss, parameters after the optimizer steps, final memory.
Output: tests/golden/sequence_golden.pt (small, committed).
"""
import os
import pickle
import sys
import tarfile
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyg_shim  # noqa: E402

REF = pyg_shim.REFERENCE_ROOT
L, N_SEQ, META_B, K = 7, 16, 8, 3
MODEL_KW = dict(edge_dim=2, memory_dim=26, time_dim=1, node_dim=2, num_iters=K, epsilon=0.5, gamma=0.1,
                readout_hidden=13, readout_out=1, init_time=0, final_act="tanh", predict_dst=True)


class _Bag:
    def __setstate__(self, st):
        self.__dict__["_st"] = st


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("torch_geometric"):
            return type(name, (_Bag,), {})
        return super().find_class(module, name)


_pm = types.ModuleType("_seq_pickle")
_pm.Unpickler = _Unpickler
_pm.__name__ = "pickle"
for _k in ("load", "loads", "dump", "dumps", "Pickler", "UnpicklingError", "HIGHEST_PROTOCOL"):
    setattr(_pm, _k, getattr(pickle, _k))


def _tensors(obj):
    st = obj._st
    while not (isinstance(st, dict) and "src" in st):
        if isinstance(st, dict):
            st = st.get("_store", st.get("_mapping", next(iter(st.values()))))
        elif isinstance(st, (tuple, list)):
            st = next(s for s in st if s is not None)
        else:
            st = getattr(st, "_st", None) or st.__dict__
    return st


def _import_reference_loop():
    ns = pyg_shim.install()
    for name in ("ray", "wandb", "clint", "clint.textui", "matplotlib", "matplotlib.pyplot", "tgb", "tgb.linkproppred",
                 "tgb.linkproppred.dataset_pyg", "torch_geometric.datasets", "tgb.linkproppred.evaluate"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["ray"].remote = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
    sys.modules["tgb.linkproppred.dataset_pyg"].PyGLinkPropPredDataset = object
    sys.modules["tgb.linkproppred.evaluate"].Evaluator = object
    sys.modules["torch_geometric.datasets"].JODIEDataset = object
    sys.modules["models"] = ns.models
    sys.path.insert(0, REF)
    import train_sequence
    import utils as ref_utils
    return ns, train_sequence, ref_utils


def main():
    ns, ts, ref_utils = _import_reference_loop()
    tmp = tempfile.mkdtemp()
    with tarfile.open(os.path.join(REF, "data", "label_sequence.tar.xz")) as tf:
        tf.extractall(tmp)
    stats, seqs = {}, None
    for dp, _, files in os.walk(tmp):
        for f in files:
            if f.startswith("._"):
                continue
            p = os.path.join(dp, f)
            if f == "delta_t_stats.pkl":
                length = int(p.split("labelsequence_")[1].split("_")[0])
                with open(p, "rb") as fh:
                    st = pickle.load(fh)
                stats[length] = st
            elif f.endswith(".pt") and f"seq_len{L}_" in f:
                seqs = torch.load(p, pickle_module=_pm, weights_only=False)
    print("known answers:", stats)
    data = []
    for obj in seqs[:N_SEQ]:
        m = _tensors(obj)
        data.append(ns.TemporalData(src=m["src"], dst=m["dst"], t=m["t"], msg=m["msg"], y=m["y"],
                                    x=m["x"][:, :N_SEQ * L].clone()))
    num_nodes = N_SEQ * L
    s0 = stats[L]
    mean, std = (s0["mean_delta_t"], s0["std_delta_t"]) if isinstance(s0, dict) else s0[:2]
    torch.manual_seed(0)
    model = ns.CTAN(num_nodes=num_nodes, mean_delta_t=mean, std_delta_t=std, **MODEL_KW)
    init_sd = {k: v.clone() for k, v in model.state_dict().items()}
    opt = torch.optim.Adam(model.parameters(), lr=3e-3, weight_decay=1e-5)
    crit = torch.nn.BCEWithLogitsLoss()
    loader = ns.LastNeighborLoader(num_nodes, size=5)
    helper = torch.empty(num_nodes, dtype=torch.long)
    meta = [torch.arange(i, i + META_B) for i in range(0, N_SEQ, META_B)]
    loaders = [ns.TemporalDataLoader(d, batch_size=1) for d in data]
    with pyg_shim.stable_sort():
        ts.train(data, model, opt, meta, loaders, crit, loader, helper, device="cpu")
        after = {k: v.clone() for k, v in model.state_dict().items()}
        out = ts.eval(data, model, meta, loaders, crit, loader, helper, return_predictions=True, device="cpu")
    scores, (y_true, _, y_pred) = out          # y_pred: the raw logits of each sequence's last event
    print({k: float(v) for k, v in scores.items() if k in ("loss", "accuracy")})
    gold = {
        "L": L, "n_seq": N_SEQ, "meta_batch": META_B, "model_kw": dict(MODEL_KW, num_nodes=num_nodes,
                                                                      mean_delta_t=float(mean), std_delta_t=float(std)),
        "lr": 3e-3, "weight_decay": 1e-5, "split": (0.15, 0.15),
        "stats": {int(k): ((float(v["mean_delta_t"]), float(v["std_delta_t"])) if isinstance(v, dict) else
                           (float(v[0]), float(v[1]))) for k, v in stats.items()},
        "msg": torch.stack([d.msg for d in data]), "y": torch.cat([d.y for d in data]), "x": data[0].x.clone(),
        "init": init_sd, "after_train": after, "eval_loss": float(scores["loss"]),
        "eval_pred": torch.as_tensor(y_pred).float().view(-1), "eval_true": torch.as_tensor(y_true).float().view(-1),
        "memory_after_eval": model.memory.memory.clone(), "last_update_after_eval": model.memory.last_update.clone(),
    }
    torch.save(gold, os.path.join(HERE, "sequence_golden.pt"))
    print("wrote sequence_golden.pt", os.path.getsize(os.path.join(HERE, "sequence_golden.pt")), "bytes")


if __name__ == "__main__":
    main()
