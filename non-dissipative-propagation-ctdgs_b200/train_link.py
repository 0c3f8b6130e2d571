"""The reference's loop functions, same signatures and return values, on the zero-host-sync engines.

    tgb_train(data, model, optimizer, train_loader, criterion, neighbor_loader, helper, device, pbar)   train_link.py:206
    tgb_test(data, model, loader, neg_sampler, neighbor_loader, split_mode, helper, evaluator, metric, ...) :24
    train(data, model, optimizer, train_loader, criterion, neighbor_loader, helper, train_neg_sampler, ...)   :287
    eval(data, model, loader, criterion, neighbor_loader, helper, neg_sampler, eval_seed, ...)               :349

`main.py` / `main_tgb_ctan.py` change ONE import (`from ctan_b200.train_link import tgb_train, tgb_test, train,
eval`, models from `ctan_b200.models`) and keep their drivers.  What differs from the reference loop is only HOW an
epoch is executed:

  * the reference slices a batch on the host, draws negatives, runs ~550 small ops and calls `loss.item()` per
    batch; here the epoch's negatives are drawn first (same RNG calls, same order), then ALL full batches run as
    CUDA-graph replays of `StepEngine` / `TGBEvalEngine` with no host synchronisation, and the per-batch losses /
    scores are read back once at the end;
  * the trailing short batch (the loaders do not drop it) and every configuration the engines do not cover
    (ragged negative sets, `val_fast`, exotic criteria, events that are not a contiguous slice of `data`) go through
    `_module_epoch`: the reference's per-batch procedure on the same modules (still our kernels, never a CPU path);
  * `optimizer` stays the owner of the Adam state (StepEngine.attach_optimizer shares storage with it), so the
    reference's checkpoint / resume code (train_link.py:497-511, 587-603) works unchanged;
  * `helper` is accepted for signature parity; the engines keep their global->local map inside the loader object.

State (memory, last_update, neighbour tables, parameters, optimizer moments) lives in the caller's objects exactly
as in the reference: the functions can be mixed freely with hand-written loops over the same objects."""
from __future__ import annotations

import datetime
import time
import types

import numpy as np
import torch

from . import _lib
from .engine import LOSS_MODES, StepEngine
from .modules import CTAN_tgb
from .negatives import PackedNegatives
from .tgb_eval import TGBEvalEngine


# ----------------------------------------------------------------------------------------------- data containers
class TemporalData:
    """Attribute bag with torch_geometric.data.TemporalData's slicing rule: indexing applies
    to every tensor whose first dimension equals num_events and leaves the rest (e.g. `x`) untouched.  PyG's own
    class works with the loop functions too; this one exists so that they run where PyG is not installed."""

    def __init__(self, **tensors):
        self.__dict__.update(tensors)

    @property
    def num_events(self):
        return self.src.size(0)

    @property
    def num_nodes(self):
        return int(max(int(self.src.max()), int(self.dst.max()))) + 1

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def __contains__(self, key):
        return key in self.__dict__

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return getattr(self, idx)
        n, out = self.num_events, TemporalData()
        for k, v in self.__dict__.items():
            out.__dict__[k] = v[idx] if (torch.is_tensor(v) and v.dim() > 0 and v.size(0) == n) else v
        return out

    def to(self, device):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device)
        return self

    def train_val_test_split(self, val_ratio: float = 0.15, test_ratio: float = 0.15):
        """Chronological split used by link_prediction_single (train_link.py:689): cuts at the quantiles of t."""
        q = np.quantile(self.t.cpu().numpy(), [1.0 - val_ratio - test_ratio, 1.0 - test_ratio])
        a, b = int((self.t <= q[0]).sum()), int((self.t <= q[1]).sum())
        return self[:a], self[a:b], self[b:]


class TemporalDataLoader:
    """Sequential mini-batches `data[i:i+batch_size]`, last one short (torch_geometric.loader.TemporalDataLoader)."""

    def __init__(self, data, batch_size: int = 1, **_):
        self.data, self.batch_size = data, int(batch_size)

    def __len__(self):
        return (self.data.num_events + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for i in range(0, self.data.num_events, self.batch_size):
            yield self.data[i:i + self.batch_size]


# ----------------------------------------------------------------------------------------------- plumbing
def _split_of(loader):
    """(events of the loader as one TemporalData-like object, batch size)."""
    d, bs = getattr(loader, "data", None), getattr(loader, "batch_size", None)
    if d is not None and bs:
        return d, int(bs)
    batches = list(loader)
    if not batches:
        return None, 0
    bs = int(batches[0].src.numel())
    keys = [k for k in ("src", "dst", "t", "msg", "y") if hasattr(batches[0], k)]
    d = TemporalData(**{k: torch.cat([getattr(b, k) for b in batches]) for k in keys})
    if hasattr(batches[0], "x"):
        d.x = batches[0].x
    return d, bs


def _aligned_start(data, split, start: int) -> bool:
    """True when `split`'s events are rows [start, start+n) of `data` -- the layout every reference driver
    produces (chronological prefix / suffix splits; the loader's event ids index `data.msg`, train_link.py:265)."""
    n = split.src.numel()
    if n == 0 or start < 0 or start + n > data.src.numel():
        return False
    for k in ("src", "dst", "t"):
        a, b = getattr(split, k), getattr(data, k)
        if a.data_ptr() == b.data_ptr() + start * a.element_size():
            continue
        if not (torch.equal(a[:1].to(b.device), b[start:start + 1])
                and torch.equal(a[-1:].to(b.device), b[start + n - 1:start + n])):
            return False
    return True


def _node_features(data, split):
    for obj in (split, data):
        x = getattr(obj, "x", None)
        if x is not None:
            return x
    return None


def _is_plain_bce(criterion) -> bool:
    return (type(criterion).__name__ == "BCEWithLogitsLoss" and criterion.reduction == "mean"
            and criterion.weight is None and criterion.pos_weight is None)


def _engines(model) -> dict:
    return model.__dict__.setdefault("_ctan_engines", {})


def _step_engine(model, neighbor_loader, optimizer, data, split, B, n_neg, criterion, n_batches, start, dev, out=False):
    """StepEngine over the FULL event table `data` (event ids index it), cached on the model."""
    x = _node_features(data, split)
    cname = type(criterion).__name__ if criterion is not None else None
    key = ("step", id(neighbor_loader), B, n_neg, data.src.data_ptr(), data.msg.data_ptr(), int(data.src.numel()),
           cname, x.data_ptr() if x is not None else 0, bool(out))
    cache = _engines(model)
    ent = cache.get(key)
    n_split = int(split.src.numel())
    if ent is not None and (ent["cap"] < n_batches or ent["neg_rows"] < n_split or ent["start"] != start):
        ent = None                                           # another split through the same tables: rebuild
    if ent is None:
        st = {"src": data.src.to(dev), "dst": data.dst.to(dev), "t": data.t.to(dev), "msg": data.msg.to(dev), "x": x}
        neg = None
        if n_neg:
            neg = torch.zeros((n_split, n_neg), dtype=torch.long, device=dev)
            st["neg"] = neg
        else:
            st["y"] = data.y.to(dev)
        eng = StepEngine(model, neighbor_loader, st, batch_size=B, n_neg=n_neg, hist_cap=max(n_batches + 1, 64),
                         criterion=criterion if n_neg == 0 else None, out_cap=(n_batches + 1) if out else 0,
                         neg_base=start)
        keep = (model.memory.memory.clone(), model.memory.last_update.clone(), neighbor_loader.e_id.clone(),
                neighbor_loader.neighbors.clone(), neighbor_loader._deg.clone(), neighbor_loader.cur_e_id)
        eng.warmup()                                         # throw-away step on scratch state, then restore
        model.memory.memory.copy_(keep[0]); model.memory.last_update.copy_(keep[1])
        neighbor_loader.e_id.copy_(keep[2]); neighbor_loader.neighbors.copy_(keep[3]); neighbor_loader._deg.copy_(keep[4])
        neighbor_loader.cur_e_id = keep[5]
        ent = {"eng": eng, "neg": neg, "cap": n_batches, "neg_rows": n_split, "start": start}
        cache[key] = ent
    eng = ent["eng"]
    if optimizer is not None:
        if eng._opt is not optimizer:
            eng.attach_optimizer(optimizer)
        else:
            eng.pull_optimizer()
    return eng, ent["neg"]


class _Batch(types.SimpleNamespace):
    def to(self, device):
        return self


def _module_epoch(data, model, optimizer, batches, criterion, neighbor_loader, helper, draw_neg, tgb, train_mode,
                  device, labelled=False):
    """The reference's per-batch procedure (train_link.py:237-282 / 297-343 / 361-405) on our modules, for the
    batches the engines do not cover.  Returns per-batch (loss, pos_out, neg_out) with detached outputs."""
    out = []
    for batch in batches:
        src, pos_dst, t, msg = batch.src.to(device), batch.dst.to(device), batch.t.to(device), batch.msg.to(device)
        if train_mode and optimizer is not None:
            optimizer.zero_grad()
        neg_dst = None if labelled else draw_neg(src)
        seeds = torch.cat([src, pos_dst] + ([neg_dst] if neg_dst is not None else []))
        n_id, edge_index, e_id, _ = neighbor_loader.lookup_khop(seeds, model.num_gnn_layers)
        helper[n_id] = torch.arange(n_id.size(0), device=device)
        x = getattr(batch, "x", None)
        if x is None:
            x = getattr(data, "x", None)
        if tgb:
            b = _Batch(src=torch.cat([src, src]), dst=torch.cat([pos_dst, neg_dst]), n_neg=int(neg_dst.numel()), x=x)
        elif neg_dst is not None:
            b = _Batch(src=src, dst=pos_dst, neg_dst=neg_dst, x=x)
        else:
            b = _Batch(src=src, dst=pos_dst, x=x)
        with torch.set_grad_enabled(train_mode):
            pos_out, neg_out, src_emb, dst_emb = model(batch=b, n_id=n_id, msg=data.msg[e_id].to(device),
                                                       t=data.t[e_id].to(device), edge_index=edge_index, id_mapper=helper)
            if labelled:
                loss = criterion(pos_out, batch.y.to(device))
            else:
                loss = criterion(pos_out, torch.ones_like(pos_out)) + criterion(neg_out, torch.zeros_like(neg_out))
        model.update(src, pos_dst, t, msg, src_emb, dst_emb)
        neighbor_loader.insert(src, pos_dst)
        if train_mode and optimizer is not None:
            loss.backward()
            optimizer.step()
        model.detach_memory()
        out.append((loss.item(), pos_out.detach(), None if neg_out is None else neg_out.detach()))
    return out


def _tail(split, B, n_full):
    """The trailing short batch of a split (loaders keep it), as a one-element batch list."""
    n = split.src.numel()
    if n_full * B >= n:
        return []
    return [split[n_full * B:n]]


def _dst_range(data):
    return int(data.dst.min()), int(data.dst.max())


# ----------------------------------------------------------------------------------------------- tgb_train
def tgb_train(data, model, optimizer, train_loader, criterion, neighbor_loader, helper, device="cpu", pbar=False):
    """One training epoch over `train_loader` (train_link.py:206-284): returns the list of per-batch losses."""
    device = torch.device(device)
    model.train()
    data.msg = data.msg.to(device)
    data.t = data.t.to(device)
    model.reset_memory()
    neighbor_loader.reset_state()
    split, B = _split_of(train_loader)
    if split is None:
        return []
    lo, hi = _dst_range(data)

    def draw(src):                                           # train_link.py:245-251
        return torch.randint(lo, hi + 1, (src.size(0),), dtype=torch.long, device=device)
    n = int(split.src.numel())
    n_full = n // B
    tgb = isinstance(model, CTAN_tgb)
    if not (_is_plain_bce(criterion) and n_full > 0 and _aligned_start(data, split, 0)):
        res = _module_epoch(data, model, optimizer, list(train_loader), criterion, neighbor_loader, helper, draw, tgb,
                            True, device)
        return [r[0] for r in res]
    eng, neg = _step_engine(model, neighbor_loader, optimizer, data, split, B, 1, criterion, n_full, 0, device)
    src = split.src.to(device)
    for j in range(n_full):                                  # the epoch's negatives: same RNG calls, same order
        neg[j * B:(j + 1) * B, 0] = draw(src[j * B:(j + 1) * B])
    eng.reset()
    eng.run(n_full, train=True)
    losses = eng.losses(n_full).tolist()
    eng.check()
    eng.push_optimizer()
    tail = _module_epoch(data, model, optimizer, _tail(split, B, n_full), criterion, neighbor_loader, helper, draw,
                         tgb, True, device)
    return losses + [r[0] for r in tail]


# ----------------------------------------------------------------------------------------------- train
def train(data, model, optimizer, train_loader, criterion, neighbor_loader, helper, train_neg_sampler=None,
          requires_grad=True, device="cpu", pbar=False):
    """One epoch of the main.py path (train_link.py:287-345).  `train_neg_sampler=None` is the labelled case (link
    regression / multi-class classification: loss = criterion(pos_out, batch.y))."""
    device = torch.device(device)
    model.train()
    data.msg = data.msg.to(device)
    data.t = data.t.to(device)
    model.reset_memory()
    neighbor_loader.reset_state()
    split, B = _split_of(train_loader)
    if split is None:
        return []
    labelled = train_neg_sampler is None

    def draw(src):
        return train_neg_sampler.sample(src).to(device)
    n = int(split.src.numel())
    n_full = n // B
    ok = n_full > 0 and _aligned_start(data, split, 0) and not isinstance(model, CTAN_tgb)
    if labelled:
        ok = ok and type(criterion).__name__ in LOSS_MODES and hasattr(data, "y") and data.y.size(0) == data.src.size(0)
    else:
        ok = ok and _is_plain_bce(criterion)
    if not ok:
        res = _module_epoch(data, model, optimizer if requires_grad else None, list(train_loader), criterion,
                            neighbor_loader, helper, draw, False, bool(requires_grad), device, labelled)
        return [r[0] for r in res]
    eng, neg = _step_engine(model, neighbor_loader, optimizer if requires_grad else None, data, split, B,
                            0 if labelled else 1, criterion, n_full, 0, device)
    if not labelled:
        src = split.src.to(device)
        for j in range(n_full):                              # sampler calls in batch order, as the reference makes them
            neg[j * B:(j + 1) * B, 0] = draw(src[j * B:(j + 1) * B])
    eng.reset()
    eng.run(n_full, train=bool(requires_grad))
    losses = eng.losses(n_full).tolist()
    eng.check()
    if requires_grad:
        eng.push_optimizer()
    tail = _module_epoch(data, model, optimizer if requires_grad else None, _tail(split, B, n_full), criterion,
                         neighbor_loader, helper, draw, False, bool(requires_grad), device, labelled)
    return losses + [r[0] for r in tail]


# ----------------------------------------------------------------------------------------------- eval
@torch.no_grad()
def eval(data, model, loader, criterion, neighbor_loader, helper, neg_sampler=None, eval_seed=12345,
         regression=False, multiclass=False, return_predictions=False, device="cpu", eval_name="eval",
         wandb_log=False, pbar=False):
    """No-grad pass over `loader` continuing from the current state (train_link.py:349-432): returns
    (scores, (y_true, y_pred, y_pred_confidence) or None).  `wandb_log` is accepted and ignored (no wandb here)."""
    from .metrics import scoring
    device = torch.device(device)
    t0 = time.time()
    model.eval()
    data.msg = data.msg.to(device)
    data.t = data.t.to(device)
    split, B = _split_of(loader)
    labelled = neg_sampler is None

    def draw(src):                                           # re-seeded every batch, as upstream
        return neg_sampler.sample(src, eval=True, eval_seed=eval_seed).to(device)
    n = int(split.src.numel())
    n_full = n // B
    start = int(neighbor_loader.cur_e_id)
    ok = n_full > 0 and _aligned_start(data, split, start) and not isinstance(model, CTAN_tgb)
    if labelled:
        ok = ok and type(criterion).__name__ in LOSS_MODES and hasattr(data, "y") and data.y.size(0) == data.src.size(0)
    else:
        ok = ok and _is_plain_bce(criterion)
    outs = []                                                # per batch: (pos_out [b,O], neg_out [b,O] or None)
    if ok:
        eng, neg = _step_engine(model, neighbor_loader, None, data, split, B, 0 if labelled else 1, criterion, n_full,
                                start, device, out=True)
        if not labelled:
            src = split.src.to(device)
            for j in range(n_full):
                neg[j * B:(j + 1) * B, 0] = draw(src[j * B:(j + 1) * B])
        eng.w["counters"][3] = 0                              # score / loss slots of this pass start at 0
        eng.set_cursor(start)
        eng.run(n_full, train=False)
        o = eng.outputs(n_full)                              # [n_full, P, O]
        eng.check()
        for j in range(n_full):
            outs.append((o[j, :B], None if labelled else o[j, B:]))
        rest = _tail(split, B, n_full)
    else:
        rest = list(loader)
    for r in _module_epoch(data, model, None, rest, criterion, neighbor_loader, helper, draw, False, False, device,
                           labelled):
        outs.append((r[1], r[2]))
    y_pred_l, y_true_l, y_conf_l = [], [], []
    pos = 0
    for po, no in outs:
        b = po.size(0)
        if labelled:
            y_true = split.y[pos:pos + b].cpu()
            y_pred = po.detach().cpu()
            y_pred_l.append(torch.argmax(y_pred, -1) if multiclass else y_pred)
        else:
            y_pred = torch.cat([po, no], dim=0).cpu()
            y_true = torch.cat([torch.ones(po.size(0)), torch.zeros(no.size(0))], dim=0)
            y_pred_l.append((y_pred.sigmoid() > 0.5).float())
        y_conf_l.append(y_pred)
        y_true_l.append(y_true)
        pos += b
    t1 = time.time()
    y_true = torch.cat(y_true_l)
    if not labelled:
        y_true = y_true.unsqueeze(1)
    y_pred, y_conf = torch.cat(y_pred_l), torch.cat(y_conf_l)
    scores = scoring(y_true, y_pred, y_conf, is_regression=regression, is_multiclass=multiclass)
    scores["loss"] = criterion(y_conf, y_true).item()
    scores["time"] = datetime.timedelta(seconds=t1 - t0)
    return scores, ((y_true, y_pred, y_conf) if return_predictions else None)


# ----------------------------------------------------------------------------------------------- tgb_test
def _packed_negatives(neg_sampler, split, split_mode, device):
    """The split's negatives as a device CSR (packed once per (sampler, split) and cached on the sampler)."""
    if isinstance(neg_sampler, PackedNegatives):
        return neg_sampler
    cache = neg_sampler.__dict__.setdefault("_ctan_packed", {})
    key = (split_mode, int(split.src.numel()), int(split.t[0]), int(split.t[-1]))
    if key not in cache:
        es = getattr(neg_sampler, "eval_set", None)
        if isinstance(es, dict) and es.get(split_mode) is not None:
            cache[key] = PackedNegatives.from_eval_set(es[split_mode], split.src, split.dst, split.t, device=device)
        else:                                                # any object with the reference's query_batch()
            rows = neg_sampler.query_batch(split.src, split.dst, split.t, split_mode=split_mode)
            counts = torch.tensor([0] + [len(r) for r in rows], dtype=torch.long)
            flat = torch.tensor([v for r in rows for v in r], dtype=torch.long)
            cache[key] = PackedNegatives(torch.cumsum(counts, 0).to(device), flat.to(device))
    return cache[key]


@torch.no_grad()
def tgb_test(data, model, loader, neg_sampler, neighbor_loader, split_mode, helper, evaluator, metric,
             validation_subsample=None, batched_val=False, device="cpu", pbar=False):
    """TGB evaluation pass (train_link.py:24-203): returns ({"mrr": mean over the pass}, None).
    'val' / 'test': one-vs-many ranking against the sampler's negatives (TGBEvalEngine: one forward per batch instead
    of one per query, MRR on the device);  'train': inference over the split with random negatives, each batch ranked
    the way Evaluator.eval does for 1-D negatives;  'val_fast': the reference's per-query procedure on our modules."""
    device = torch.device(device)
    model.eval()
    if metric != "mrr":
        raise NotImplementedError("tgb_test computes the MRR (the metric of every TGB link dataset)")
    split, B = _split_of(loader)
    n = int(split.src.numel())
    n_full = n // B
    start = int(neighbor_loader.cur_e_id)
    aligned = n_full > 0 and _aligned_start(data, split, start) and isinstance(model, CTAN_tgb)
    lo, hi = _dst_range(data)
    perf = []                                                # one entry per query (val/test) or per batch (train)

    def draw(src):
        return torch.randint(lo, hi + 1, (src.size(0),), dtype=torch.long, device=device)

    if split_mode == "train":
        crit = torch.nn.BCEWithLogitsLoss()
        outs = []
        if aligned:
            eng, neg = _step_engine(model, neighbor_loader, None, data, split, B, 1, crit, n_full, start, device, out=True)
            src = split.src.to(device)
            for j in range(n_full):
                neg[j * B:(j + 1) * B, 0] = draw(src[j * B:(j + 1) * B])
            eng.w["counters"][3] = 0
            eng.set_cursor(start)
            eng.run(n_full, train=False)
            o = eng.outputs(n_full)[..., 0]                  # [n_full, 2B]
            eng.check()
            pos_s, neg_s = o[:, :B], o[:, B:]
            # Evaluator.eval with 1-D negatives ranks every positive of the batch against all its negatives
            ge = (neg_s[:, None, :] >= pos_s[:, :, None]).sum(-1)
            gt = (neg_s[:, None, :] > pos_s[:, :, None]).sum(-1)
            perf += (1.0 / (0.5 * (ge + gt).float() + 1.0)).mean(1).tolist()
            rest = _tail(split, B, n_full)
        else:
            rest = list(loader)
        for _, po, no in _module_epoch(data, model, None, rest, crit, neighbor_loader, helper, draw, True, False, device):
            p_, n_ = po.view(-1, 1), no.view(1, -1)
            rank = 0.5 * ((n_ >= p_).sum(1) + (n_ > p_).sum(1)).float() + 1.0
            perf.append(float((1.0 / rank).mean()))
        return {"mrr": float(torch.tensor(perf).mean())}, None

    if split_mode in ("val", "test"):
        packed = _packed_negatives(neg_sampler, split, split_mode, device)
        if aligned and packed.is_uniform() and packed.num_events == n:
            dense = packed.dense().to(device)
            n_neg = int(dense.size(1))
            x = _node_features(data, split)
            key = ("eval", id(neighbor_loader), B, n_neg, data.src.data_ptr(), data.msg.data_ptr(), dense.data_ptr(), start)
            cache = _engines(model)
            if key not in cache:
                st = {"src": data.src.to(device), "dst": data.dst.to(device), "t": data.t.to(device),
                      "msg": data.msg.to(device), "neg": dense, "x": x}
                ev = TGBEvalEngine(model, neighbor_loader, st, batch_size=B, n_neg=n_neg, neg_base=start)
                ev.warmup()
                cache[key] = ev
            ev = cache[key]
            ev.reset_metric()
            ev.set_cursor(start)
            ev.run(n_full)
            ev.check()
            acc = ev.w["acc"].cpu()
            total, count = float(acc[0]), float(acc[1])
            for r in _per_query_eval(data, model, _tail(split, B, n_full), packed, n_full * B, neighbor_loader, helper,
                                     device, fast=False):
                total += r
                count += 1
            return {"mrr": total / max(count, 1.0)}, None
        perf = _per_query_eval(data, model, list(loader), packed, 0, neighbor_loader, helper, device, fast=False)
        return {"mrr": float(np.mean(perf))}, None

    if split_mode == "val_fast":
        packed = _packed_negatives(neg_sampler, split, "val", device)
        perf = _per_query_eval(data, model, list(loader), packed, 0, neighbor_loader, helper, device, fast=True,
                               subsample=validation_subsample)
        return {"mrr": float(np.mean(perf))}, None
    raise ValueError(f"unknown split_mode {split_mode!r}")


def _per_query_eval(data, model, batches, packed, first_event, neighbor_loader, helper, device, fast, subsample=None):
    """The reference's one-forward-per-query evaluation (train_link.py:60-159) on our modules; used for what the
    batch engine does not cover.  Returns the per-query reciprocal ranks."""
    rr = []
    pos = first_event
    for batch in batches:
        src, dst, t, msg = batch.src.to(device), batch.dst.to(device), batch.t.to(device), batch.msg.to(device)
        b = int(src.numel())
        rows = packed.query_batch(pos, pos + b)
        pos += b
        s_embs, d_embs = [], []
        x = getattr(batch, "x", None)
        if x is None:
            x = getattr(data, "x", None)
        for q, neg_list in enumerate(rows):
            if fast and subsample is not None:
                neg_list = np.random.choice(neg_list, size=subsample, replace=True)
            neg = torch.as_tensor(np.asarray(neg_list), dtype=torch.long, device=device)
            s_rep = src[q].repeat(1 + neg.numel())
            seeds = torch.cat([src[q].view(1), dst[q].view(1), neg]) if fast else torch.cat([s_rep, dst, neg])
            n_id, edge_index, e_id, _ = neighbor_loader.lookup_khop(seeds, model.num_gnn_layers)
            helper[n_id] = torch.arange(n_id.size(0), device=device)
            qb = _Batch(src=s_rep, dst=torch.cat([dst[q].view(1), neg]), x=x, n_neg=int(neg.numel()))
            po, no, se, de = model(batch=qb, n_id=n_id, msg=data.msg[e_id].to(device), t=data.t[e_id].to(device),
                                   edge_index=edge_index, id_mapper=helper)
            p_, n_ = po.view(-1), no.view(-1)
            rank = 0.5 * float((n_ >= p_).sum() + (n_ > p_).sum()) + 1.0
            rr.append(1.0 / rank)
            s_embs.append(se)
            d_embs.append(de)
        model.update(src, dst, t, msg, torch.cat(s_embs, 0), torch.cat(d_embs, 0))
        neighbor_loader.insert(src, dst)
    return rr
