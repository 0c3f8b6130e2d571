"""The reference's sequence-classification loops (config C1, Temporal Path Graph), same signatures, on the engine.

    train(train_data, model, optimizer, train_meta_loader, train_loaders, criterion, neighbor_loader, helper, device)
                                                                                                train_sequence.py:16
    eval(eval_data, model, meta_loader, eval_loaders, criterion, neighbor_loader, helper, return_predictions, ...)  :68

What the reference does per event (one event per step, K loader calls with `e_id % (L-1)`, forward, memory write-back,
insert) becomes one CUDA-graph replay of `StepEngine` in sequence mode.  Only the LAST event of a sequence enters the
loss and the memory is stored detached (memory_layers.py:144), so the gradient of a meta-batch is the sum over its
sequences of the last event's gradient, scaled by 1/len(idx) (criterion's mean): every sequence is L-2 no-grad steps
plus one gradient-accumulating step, and the optimizer step runs once per meta-batch -- the same arithmetic as
`loss.backward(); optimizer.step()` at train_sequence.py:58-60.

The pass is laid out on the device as ONE stream in processing order (the meta loader's order, which may be shuffled),
so the host only issues graph replays.  Configurations the engine does not cover (sequences of different lengths, a
loader batch size other than 1, a model without `predict_dst`) run the reference's per-event procedure on our modules."""
from __future__ import annotations

import datetime
import time
import types

import torch

from .engine import LOSS_MODES, StepEngine


def _order(meta_loader):
    batches = [[int(i) for i in idx] for idx in meta_loader]
    return batches


def _uniform(model, data, loaders):
    """L-1 if every sequence has the same number of events, one event per loader batch, and ONE node-feature table
    (the dataset shares `x` between its sequences, datasets/label_prediction_sequence.py:33-52; after the driver's
    per-sequence `.to(device)` they are equal copies, which is checked once per data list), else 0."""
    L1 = int(data[0].src.numel())
    for d in data:
        if int(d.src.numel()) != L1:
            return 0
    for ld in (loaders.values() if isinstance(loaders, dict) else loaders):
        if getattr(ld, "batch_size", 1) != 1:
            return 0
    cache = model.__dict__.setdefault("_ctan_seq_x_ok", {})
    key = (id(data), len(data))
    if key not in cache:
        x0 = getattr(data[0], "x", None)
        ok = True
        for d in data[1:]:
            x = getattr(d, "x", None)
            if x is x0:
                continue
            if x is None or x0 is None or x.shape != x0.shape or (x.data_ptr() != x0.data_ptr() and not torch.equal(x, x0)):
                ok = False
                break
        cache[key] = ok
    return L1 if cache[key] else 0


def _stacked(model, data, dev):
    """[n_seq, L-1, ...] device copies of the sequences (cached on the model, keyed by the list object)."""
    cache = model.__dict__.setdefault("_ctan_seq_data", {})
    key = (id(data), len(data), int(data[0].src.numel()))
    if key not in cache:
        cache[key] = {k: torch.stack([getattr(d, k) for d in data]).to(dev) for k in ("src", "dst", "t", "msg")}
        cache[key]["y"] = torch.stack([d.y.reshape(-1).float() for d in data]).to(dev)          # [n_seq, O]
    return cache[key]


def _engine(model, neighbor_loader, optimizer, criterion, data, L1, n_seq_max, dev):
    cache = model.__dict__.setdefault("_ctan_engines", {})
    O = model.readout.lin_final.weight.size(0)
    x0 = getattr(data[0], "x", None)
    key = ("seq", id(neighbor_loader), L1, type(criterion).__name__, tuple(x0.shape) if x0 is not None else None)
    ent = cache.get(key)
    if ent is not None and (ent["cap"] < n_seq_max or (x0 is not None and ent["x"] is not x0
                                                       and not torch.equal(ent["x"], x0.to(ent["x"].device)))):
        ent = None                                            # larger pass, or another node-feature table: rebuild
    if ent is None:
        n_ev = n_seq_max * L1
        Fe = data[0].msg.size(1)
        st = {"src": torch.zeros(n_ev, dtype=torch.long, device=dev), "dst": torch.zeros(n_ev, dtype=torch.long, device=dev),
              "t": torch.zeros(n_ev, dtype=torch.long, device=dev), "msg": torch.zeros((n_ev, Fe), device=dev),
              "y": torch.zeros((n_ev, O), device=dev), "x": getattr(data[0], "x", None)}
        eng = StepEngine(model, neighbor_loader, st, batch_size=1, n_neg=0, criterion=criterion, seq_mod=L1,
                         out_cap=n_ev, hist_cap=max(n_ev + 1, 64))
        keep = (model.memory.memory.clone(), model.memory.last_update.clone(), neighbor_loader.e_id.clone(),
                neighbor_loader.neighbors.clone(), neighbor_loader._deg.clone(), neighbor_loader.cur_e_id)
        eng.warmup()
        model.memory.memory.copy_(keep[0]); model.memory.last_update.copy_(keep[1])
        neighbor_loader.e_id.copy_(keep[2]); neighbor_loader.neighbors.copy_(keep[3]); neighbor_loader._deg.copy_(keep[4])
        neighbor_loader.cur_e_id = keep[5]
        ent = {"eng": eng, "st": st, "cap": n_seq_max, "x": x0}
        cache[key] = ent
    eng = ent["eng"]
    if optimizer is not None:
        if eng._opt is not optimizer:
            eng.attach_optimizer(optimizer)
        else:
            eng.pull_optimizer()
    return eng, ent["st"]


def _load_pass(eng, st, stk, order, L1):
    """Copy the pass (sequences in processing order) into the engine's stream buffers and point the cursor at row 0."""
    idx = torch.tensor(order, dtype=torch.long, device=eng.dev)
    n = idx.numel() * L1
    for k in ("src", "dst", "t"):
        st[k][:n].copy_(stk[k][idx].reshape(-1))
    st["msg"][:n].copy_(stk["msg"][idx].reshape(n, -1))
    st["y"][:n].copy_(stk["y"][idx].repeat_interleave(L1, dim=0))
    start = int(eng.loader.cur_e_id)
    eng.w["counters"][0] = 0
    eng.w["counters"][3] = 0
    eng.set_eid_offset(start)
    return n


def _module_pass(data, model, optimizer, batches, loaders, criterion, neighbor_loader, helper, device, train_mode):
    """The reference's per-event procedure on our modules (fallback for what the engine does not cover)."""
    preds_all, true_all = [], []
    for idx in batches:
        if train_mode:
            optimizer.zero_grad()
        preds, trues = [], []
        for i in idx:
            d = data[i]
            last = None
            for inter in loaders[i]:
                src, dst, t, msg = inter.src.to(device), inter.dst.to(device), inter.t.to(device), inter.msg.to(device)
                n_id = torch.cat([src, dst]).unique()
                edge_index = torch.empty((2, 0), dtype=torch.long, device=device)
                e_id = neighbor_loader.e_id[n_id]
                for _ in range(model.num_gnn_layers):
                    n_id, edge_index, e_id = neighbor_loader(n_id)
                    e_id = e_id % d.msg.shape[0]
                helper[n_id] = torch.arange(n_id.size(0), device=device)
                b = types.SimpleNamespace(src=src, dst=dst, x=getattr(inter, "x", None))
                with torch.set_grad_enabled(train_mode):
                    y_pred, _, se, de = model(batch=b, n_id=n_id, msg=d.msg.to(device)[e_id], t=d.t.to(device)[e_id],
                                              edge_index=edge_index, id_mapper=helper)
                last = (y_pred.squeeze(0), inter.y.to(device))
                model.update(src, dst, t, msg, se, de)
                neighbor_loader.insert(src, dst)
            preds.append(last[0])
            trues.append(last[1])
        if train_mode:
            loss = criterion(torch.cat(preds, dim=-1), torch.cat(trues, dim=-1))
            loss.backward()
            optimizer.step()
            model.detach_memory()
        preds_all += [p.detach() for p in preds]
        true_all += trues
    return preds_all, true_all


def train(train_data, model, optimizer, train_meta_loader, train_loaders, criterion, neighbor_loader, helper,
          device="cpu"):
    """One training pass (train_sequence.py:16-65).  Returns None, like the reference."""
    device = torch.device(device)
    model.train()
    model.reset_memory()
    neighbor_loader.reset_state()
    batches = _order(train_meta_loader)
    L1 = _uniform(model, train_data, train_loaders)
    ok = (L1 >= 1 and getattr(model, "predict_dst", False) and type(criterion).__name__ in LOSS_MODES
          and LOSS_MODES[type(criterion).__name__] != 1)
    if not ok:
        _module_pass(train_data, model, optimizer, batches, train_loaders, criterion, neighbor_loader, helper, device, True)
        return
    eng, st = _engine(model, neighbor_loader, optimizer, criterion, train_data, L1, len(train_data), device)
    eng.reset()
    order = [i for b in batches for i in b]
    _load_pass(eng, st, _stacked(model, train_data, device), order, L1)
    eng.zero_grads()                                         # optimizer.zero_grad() (train_sequence.py:24); adam() re-clears
    for b in batches:
        eng.set_grad_scale(1.0 / len(b))
        for _ in b:
            for _e in range(L1 - 1):
                eng.step("infer")
            eng.step("grad")
        eng.adam()
    eng.check()
    eng.push_optimizer()


@torch.no_grad()
def eval(eval_data, model, meta_loader, eval_loaders, criterion, neighbor_loader, helper, return_predictions=False,
         eval_name="eval", is_regression=False, wandb_log=False, device="cpu"):
    """No-grad pass continuing from the current state (train_sequence.py:68-132): (scores, (y_true, y_pred, y_conf) or
    None).  `wandb_log` is accepted and ignored."""
    from .metrics import scoring
    device = torch.device(device)
    t0 = time.time()
    model.eval()
    batches = _order(meta_loader)
    order = [i for b in batches for i in b]
    L1 = _uniform(model, eval_data, eval_loaders)
    ok = (L1 >= 1 and getattr(model, "predict_dst", False) and type(criterion).__name__ in LOSS_MODES
          and LOSS_MODES[type(criterion).__name__] != 1)
    if ok:
        eng, st = _engine(model, neighbor_loader, None, criterion, eval_data, L1, len(eval_data), device)
        stk = _stacked(model, eval_data, device)
        n = _load_pass(eng, st, stk, order, L1)
        for _ in range(n):
            eng.step("infer")
        out = eng.outputs(n)                                 # [n, 1, O]
        eng.check()
        y_conf = out[L1 - 1::L1, 0, :].reshape(-1).cpu()     # the prediction at the last event of every sequence
        y_true = stk["y"][torch.tensor(order, device=device)].reshape(-1).cpu()
    else:
        preds, trues = _module_pass(eval_data, model, None, batches, eval_loaders, criterion, neighbor_loader, helper,
                                    device, False)
        y_conf, y_true = torch.cat(preds).cpu(), torch.cat(trues).cpu()
    t1 = time.time()
    y_pred = (y_conf.sigmoid() > 0.5).float()
    scores = scoring(y_true, y_pred, y_conf, is_regression=is_regression)
    scores["loss"] = criterion(y_conf, y_true).item()
    scores["time"] = datetime.timedelta(seconds=t1 - t0)
    return scores, ((y_true, y_pred, y_conf) if return_predictions else None)
