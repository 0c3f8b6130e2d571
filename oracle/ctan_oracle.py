"""CPU oracle for CTAN's per-event-batch hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs may import this
module, and only as the checker / the timed baseline.  The product (``ctan_b200``) never does.

This is a from-scratch RESTATEMENT (plain PyTorch ops + numpy) of the algorithm that the
reference executes on this path; each function cites the reference lines it follows
(paths relative to /root/reference).  The arithmetic that the reference delegates to
torch-geometric 2.3.1 / torch-scatter 2.1.2 (un-vendored, pinned in `env.yml:100-101`) is
restated from the published algorithm.

PARITY PINNING.  The reference ships no tests or golden vectors for this path,
so the oracle is pinned against outputs of the reference itself, produced in the build
container by running the reference's own `models/*.py` and `TGB/modules/*.py` unmodified
(through `oracle/pyg_shim.py`): see `tests/golden/make_golden.py` (committed fixtures under
`tests/golden/`) and `tests/test_oracle_vs_reference.py` (live cross-check when
/root/reference is present).  The one fixture the reference does ship (`data/label_sequence.tar.xz`: the
path-graph datasets and the known answers of `utils.compute_stats`) pins `delta_t_stats_ref` and, through the
reference's own `train_sequence.train/eval`, the sequence loop (`tests/golden/make_sequence_golden.py`).
The PyG `TransformerConv`/`AntiSymmetricConv`/`softmax`
operators themselves are "parity unpinned": no PyG source or wheel is available offline, so
both the shim and this file restate them from the same published description.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor


# =============================================================================
# 1. Last-neighbour ring buffer  (TGB/modules/neighbor_loader.py:15-85)
# =============================================================================
class NeighborTableSpec:
    """Literal, loop-based statement of the loader's CPU semantics.
    Pure Python/numpy: for small cases only.  State: `neighbors`, `e_id` [num_nodes,S] int64.

    insert (neighbor_loader.py:41-81): the 2B entries (dst<-src first, then src<-dst, batch
    order) are grouped by node with a STABLE sort; entry at sorted position p goes to dense
    slot p % S (last write wins) so for a node with c entries the last min(c,S) survive; the
    old row and the new dense row are merged by top-S on e_id (descending, -1 padding last).
    """

    def __init__(self, num_nodes: int, size: int):
        self.size = size
        self.neighbors = np.zeros((num_nodes, size), dtype=np.int64)
        self.e_id = np.full((num_nodes, size), -1, dtype=np.int64)
        self.cur_e_id = 0

    def reset_state(self):
        self.cur_e_id = 0
        self.e_id.fill(-1)

    def insert(self, src, dst):
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        B, S = src.shape[0], self.size
        per_node = {}
        for i in range(B):                                   # entries with node = dst[i]
            per_node.setdefault(int(dst[i]), []).append((self.cur_e_id + i, int(src[i])))
        tail = {}
        for i in range(B):                                   # entries with node = src[i]
            tail.setdefault(int(src[i]), []).append((self.cur_e_id + i, int(dst[i])))
        for u, lst in tail.items():
            per_node.setdefault(u, []).extend(lst)
        self.cur_e_id += B
        for u, lst in per_node.items():
            new = lst[-S:]                                   # last min(c,S) in stable order survive
            new = sorted(new, key=lambda p: -p[0])           # top-k => descending e_id
            old = [(int(self.e_id[u, s]), int(self.neighbors[u, s])) for s in range(S) if self.e_id[u, s] >= 0]
            row = (new + old)[:S]                            # every new e_id exceeds every old one
            for s in range(S):
                if s < len(row):
                    self.e_id[u, s], self.neighbors[u, s] = row[s]
                else:
                    self.e_id[u, s] = -1

    def lookup(self, n_id):
        """neighbor_loader.py:25-39.  Returns (n_id', edge_index[2,E], e_id[E])."""
        n_id = np.asarray(n_id, dtype=np.int64)
        nb, ct, ev = [], [], []
        for u in n_id:
            for s in range(self.size):
                if self.e_id[u, s] >= 0:
                    nb.append(self.neighbors[u, s]); ct.append(u); ev.append(self.e_id[u, s])
        nb = np.asarray(nb, dtype=np.int64); ct = np.asarray(ct, dtype=np.int64)
        out = np.unique(np.concatenate([n_id, nb]))
        assoc = {int(g): k for k, g in enumerate(out)}
        ei = np.array([[assoc[int(a)] for a in nb], [assoc[int(c)] for c in ct]], dtype=np.int64).reshape(2, -1)
        return out, ei, np.asarray(ev, dtype=np.int64)


class TorchNeighborLoader:
    """Op-by-op torch restatement of the loader (same cost profile as the reference: sort,
    unique, dense scatter, top-k; device-agnostic so it also serves as the eager-GPU baseline).
    Follows TGB/modules/neighbor_loader.py:16-85."""

    def __init__(self, num_nodes: int, size: int, device=None):
        self.size = size
        self.neighbors = torch.zeros((num_nodes, size), dtype=torch.long, device=device)
        self.e_id = torch.empty((num_nodes, size), dtype=torch.long, device=device)
        self._assoc = torch.empty(num_nodes, dtype=torch.long, device=device)
        self.reset_state()

    def reset_state(self):
        self.cur_e_id = 0
        self.e_id.fill_(-1)

    def __call__(self, n_id: Tensor):
        S = self.size
        rows_e = self.e_id[n_id]
        rows_n = self.neighbors[n_id]
        centers = n_id.view(-1, 1).expand(-1, S)
        keep = rows_e >= 0
        e_sel, nb_sel, ct_sel = rows_e[keep], rows_n[keep], centers[keep]
        out = torch.cat([n_id, nb_sel]).unique()
        self._assoc[out] = torch.arange(out.numel(), device=out.device)
        return out, torch.stack([self._assoc[nb_sel], self._assoc[ct_sel]]), e_sel

    def insert(self, src: Tensor, dst: Tensor):
        S, B = self.size, src.numel()
        ev = torch.arange(self.cur_e_id, self.cur_e_id + B, device=src.device).repeat(2)
        self.cur_e_id += B
        node = torch.cat([dst, src])
        other = torch.cat([src, dst])
        node, order = node.sort(stable=True)
        other, ev = other[order], ev[order]
        uniq = node.unique()
        self._assoc[uniq] = torch.arange(uniq.numel(), device=uniq.device)
        slot = torch.arange(node.numel(), device=node.device) % S + self._assoc[node] * S
        fresh_e = ev.new_full((uniq.numel() * S,), -1)
        fresh_n = ev.new_zeros(uniq.numel() * S)
        # duplicate slots: the LAST write must win (CPU index_put_ order); make it explicit
        last = torch.zeros_like(fresh_e).scatter_reduce_(
            0, slot, torch.arange(node.numel(), device=node.device), reduce="amax", include_self=False)
        wins = last[slot] == torch.arange(node.numel(), device=node.device)
        fresh_e[slot[wins]] = ev[wins]
        fresh_n[slot[wins]] = other[wins]
        all_e = torch.cat([self.e_id[uniq], fresh_e.view(-1, S)], dim=1)
        all_n = torch.cat([self.neighbors[uniq], fresh_n.view(-1, S)], dim=1)
        top_e, pick = all_e.topk(S, dim=1)
        self.e_id[uniq] = top_e
        self.neighbors[uniq] = all_n.gather(1, pick)


# =============================================================================
# 2. Memory write-back  (models/memory_layers.py:132-144 + TGB/modules/msg_agg.py:15-21)
# =============================================================================
def memory_update_ref(memory: Tensor, last_update: Tensor, src: Tensor, dst: Tensor, t: Tensor,
                      src_emb: Tensor, dst_emb: Tensor) -> None:
    """For every unique node u of cat(src,dst): last_update[u] = max t, memory[u] = emb row at
    the FIRST position attaining that max (torch-scatter CPU arg-max: strict '>' update).
    In place, detached (memory_layers.py:143-144)."""
    idx = torch.cat([src, dst])
    tt = torch.cat([t, t])
    emb = torch.cat([src_emb, dst_emb]).detach()
    uniq, inv = idx.unique(return_inverse=True)
    tmax = torch.full((uniq.numel(),), torch.iinfo(torch.long).min, dtype=torch.long, device=idx.device)
    tmax = tmax.scatter_reduce(0, inv, tt, reduce="amax", include_self=True)
    pos = torch.arange(idx.numel(), device=idx.device)
    cand = torch.where(tt == tmax[inv], pos, torch.full_like(pos, idx.numel()))
    first = torch.full((uniq.numel(),), idx.numel(), dtype=torch.long, device=idx.device)
    first = first.scatter_reduce(0, inv, cand, reduce="amin", include_self=True)
    last_update[uniq] = tmax
    memory[uniq] = emb[first]


# =============================================================================
# 3. Embedding: time encoding + K x (TransformerConv attention, anti-symmetric update)
#    models/gnn_layers.py:93-98, TGB/modules/time_enc.py:23-24, PyG 2.3.1 (App. A.5)
# =============================================================================
def segment_softmax_ref(score: Tensor, seg: Tensor, n: int) -> Tensor:
    """PyG utils.softmax: exp(s - max_seg) / (sum_seg + 1e-16)."""
    mx = score.new_full((n,), -float("inf")).scatter_reduce(0, seg, score.detach(), reduce="amax", include_self=True)
    ex = (score - mx[seg]).exp()
    den = score.new_zeros(n).scatter_add_(0, seg, ex) + 1e-16
    return ex / den[seg]


def ctan_embed_ref(P: dict, z0: Tensor, last_update: Tensor, edge_index: Tensor, t: Tensor, msg: Tensor,
                   *, num_iters: int, epsilon: float, gamma: float, mean_delta_t: float, std_delta_t: float,
                   hoist_edge_proj: bool = False) -> Tensor:
    """CTANEmbedding.forward (gnn_layers.py:93-98) with AntiSymmetricConv(TransformerConv) inlined.
    `P` maps state-dict keys (prefix 'gnn.') to tensors.  The reference recomputes lin_edge every
    iteration; `hoist_edge_proj` only changes cost, not values."""
    j, i = edge_index[0], edge_index[1]                         # j: source/neighbour, i: target/centre
    N, D = z0.size(0), P["gnn.aconv.W"].size(0)
    rel = (last_update[j] - t).abs()                            # int64
    rel = ((rel - mean_delta_t) / std_delta_t).to(z0.dtype)     # int64 -> fp32 at the subtraction
    x = F.linear(z0, P["gnn.enc_x.weight"], P["gnn.enc_x.bias"])
    tenc = F.linear(rel.view(-1, 1), P["gnn.time_enc.lin.weight"], P["gnn.time_enc.lin.bias"]).cos()
    edge_attr = torch.cat([msg, tenc], dim=-1)                  # msg first, time encoding last
    W = P["gnn.aconv.W"]
    A = W - W.t() - gamma * torch.eye(D, dtype=W.dtype, device=W.device)
    e_h = F.linear(edge_attr, P["gnn.aconv.phi.lin_edge.weight"]) if hoist_edge_proj else None
    for _ in range(num_iters):
        q = F.linear(x, P["gnn.aconv.phi.lin_query.weight"], P["gnn.aconv.phi.lin_query.bias"])
        k = F.linear(x, P["gnn.aconv.phi.lin_key.weight"], P["gnn.aconv.phi.lin_key.bias"])
        v = F.linear(x, P["gnn.aconv.phi.lin_value.weight"], P["gnn.aconv.phi.lin_value.bias"])
        e = e_h if hoist_edge_proj else F.linear(edge_attr, P["gnn.aconv.phi.lin_edge.weight"])
        kj = k[j] + e
        score = (q[i] * kj).sum(-1) / math.sqrt(D)
        alpha = segment_softmax_ref(score, i, N)
        m = (v[j] + e) * alpha.view(-1, 1)
        phi = x.new_zeros(N, D).index_add_(0, i, m)
        h = x @ A.t() + phi + P["gnn.aconv.bias"]
        x = x + epsilon * torch.tanh(h)
    return x


def readout_ref(P: dict, z_src: Tensor, z_dst: Tensor) -> Tensor:
    """LinkPredictor / TGBLinkPredictor.forward (models/predictors.py:16-19, 31-34)."""
    h = F.linear(z_src, P["readout.lin_src.weight"], P["readout.lin_src.bias"]) + \
        F.linear(z_dst, P["readout.lin_dst.weight"], P["readout.lin_dst.bias"])
    return F.linear(h.relu(), P["readout.lin_final.weight"], P["readout.lin_final.bias"])


def mrr_ref(y_pos: np.ndarray, y_neg: np.ndarray) -> float:
    """TGB Evaluator numpy branch (TGB/tgb/linkproppred/evaluate.py:109-120): one positive vs its
    negatives; rank = 0.5*(#neg>=pos + #neg>pos) + 1; mrr = 1/rank in float32."""
    y_pos = np.asarray(y_pos, dtype=np.float32).reshape(-1, 1)
    y_neg = np.asarray(y_neg, dtype=np.float32).reshape(y_pos.shape[0], -1)
    opt = (y_neg >= y_pos).sum(axis=1)
    pes = (y_neg > y_pos).sum(axis=1)
    rank = 0.5 * (opt + pes) + 1
    return float((1.0 / rank.astype(np.float32)).mean())


# =============================================================================
# 4. Module wrappers with the reference's constructor signatures and state-dict keys
#    models/ctdg_models.py:382-540, models/memory_layers.py:117-156 (App. A.7, A.8)
# =============================================================================
class _Lin(torch.nn.Module):
    """weight/bias holder initialised like torch.nn.Linear / PyG Linear (identical draws)."""

    def __init__(self, fan_in, fan_out, bias=True, pyg=False):
        super().__init__()
        self.pyg = pyg
        self.weight = torch.nn.Parameter(torch.empty(fan_out, fan_in))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(fan_out))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        fan = self.weight.size(1)
        if self.pyg:      # torch_geometric.nn.inits.kaiming_uniform(fan, a=sqrt(5))
            bound = math.sqrt(6 / ((1 + math.sqrt(5) ** 2) * fan))
            self.weight.data.uniform_(-bound, bound)
        else:             # torch.nn.Linear.reset_parameters
            torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            b = 1.0 / math.sqrt(fan)
            self.bias.data.uniform_(-b, b)


class _Holder(torch.nn.Module):
    pass


class OracleSimpleMemory(torch.nn.Module):
    def __init__(self, num_nodes, memory_dim, init_time=0):
        super().__init__()
        self.num_nodes, self.memory_dim, self.init_time = num_nodes, memory_dim, init_time
        self.register_buffer("memory", torch.zeros(num_nodes, memory_dim))
        self.register_buffer("last_update", torch.ones(num_nodes, dtype=torch.long) * init_time)
        self.register_buffer("_assoc", torch.zeros(num_nodes, dtype=torch.long))

    def reset_state(self):
        self.memory.zero_()
        self.last_update.fill_(1)
        self.last_update *= self.init_time

    def detach(self):
        self.memory.detach_()

    def update_state(self, src, pos_dst, t, src_emb, pos_dst_emb):
        memory_update_ref(self.memory, self.last_update, src, pos_dst, t, src_emb, pos_dst_emb)

    def forward(self, n_id):
        return self.memory[n_id], self.last_update[n_id]


class OracleCTAN(torch.nn.Module):
    """Same constructor arguments, forward contract and state-dict keys as the reference CTAN
    (models/ctdg_models.py:382-470); `tgb=True` gives CTAN_tgb (474-540)."""

    def __init__(self, num_nodes, edge_dim, memory_dim, time_dim, node_dim=0, num_iters=1, gnn_act="tanh",
                 gnn_act_kwargs=None, epsilon=0.1, gamma=0.1, readout_hidden=None, readout_out=1,
                 mean_delta_t=0.0, std_delta_t=1.0, init_time=0, final_act=None, predict_dst=False,
                 conv_type="TransformerConv", separable=False, tgb=False):
        super().__init__()
        assert conv_type == "TransformerConv" and gnn_act == "tanh"
        self.predict_dst = predict_dst
        D, T = memory_dim, time_dim
        self.num_nodes, self.num_gnn_layers, self.tgb = num_nodes, num_iters, tgb
        self.epsilon, self.gamma = epsilon, gamma
        self.mean_delta_t, self.std_delta_t = mean_delta_t, std_delta_t
        self.final_act = final_act
        self.memory = OracleSimpleMemory(num_nodes, D, init_time)
        # --- parameter draw order of App. A.7 -------------------------------------------
        gnn = _Holder()
        gnn.enc_x = _Lin(D + node_dim, D)
        phi = _Holder()
        phi.lin_key, phi.lin_query, phi.lin_value = _Lin(D, D, pyg=True), _Lin(D, D, pyg=True), _Lin(D, D, pyg=True)
        phi.lin_edge, phi.lin_skip = _Lin(edge_dim + T, D, bias=False, pyg=True), _Lin(D, D, pyg=True)
        order = [phi.lin_key, phi.lin_query, phi.lin_value, phi.lin_edge, phi.lin_skip]
        for m in order:                       # TransformerConv.__init__ ends with reset_parameters()
            m.reset_parameters()
        aconv = _Holder()
        aconv.W = torch.nn.Parameter(torch.empty(D, D))
        aconv.register_buffer("eye", torch.eye(D))
        aconv.phi = phi
        aconv.bias = torch.nn.Parameter(torch.empty(D))
        torch.nn.init.kaiming_uniform_(aconv.W, a=math.sqrt(5))
        for m in order:                       # AntiSymmetricConv.reset_parameters -> phi.reset_parameters()
            m.reset_parameters()
        aconv.bias.data.zero_()
        gnn.aconv = aconv
        te = _Holder()
        te.lin = _Lin(1, T)
        gnn.time_enc = te
        self.gnn = gnn
        H = D if readout_hidden is None else readout_hidden
        ro = _Holder()
        if predict_dst:                       # SequencePredictor (predictors.py:37-48)
            ro.lin, ro.lin_final = _Lin(D, H), _Lin(H, readout_out)
        else:
            ro.lin_src, ro.lin_dst, ro.lin_final = _Lin(D, H), _Lin(D, H), _Lin(H, readout_out)
        self.readout = ro
        if tgb:                               # CTAN_tgb replaces the readout (ctdg_models.py:502)
            ro = _Holder()
            ro.lin_src, ro.lin_dst, ro.lin_final = _Lin(D, D), _Lin(D, D), _Lin(D, 1)
            self.readout = ro

    # plumbing of GenericModel (ctdg_models.py:24-34, 435-436)
    def reset_memory(self):
        self.memory.reset_state()

    def detach_memory(self):
        self.memory.detach()

    def update(self, src, pos_dst, t, msg, src_emb, pos_dst_emb):
        self.memory.update_state(src, pos_dst, t, src_emb, pos_dst_emb)

    def _P(self):
        return dict(self.named_parameters())

    def embed(self, batch, n_id, msg, t, edge_index):
        z, last_update = self.memory(n_id)
        if hasattr(batch, "x") and batch.x is not None:
            x = batch.x.squeeze(0) if batch.x.dim() == 3 else batch.x
            if x.dim() != 2:
                raise ValueError(f"Unexpected node feature shape. Got {batch.x.shape}")
            z = torch.cat((z, x[n_id]), dim=-1)
        z = ctan_embed_ref(self._P(), z, last_update, edge_index, t, msg, num_iters=self.num_gnn_layers,
                           epsilon=self.epsilon, gamma=self.gamma, mean_delta_t=self.mean_delta_t,
                           std_delta_t=self.std_delta_t)
        if self.final_act is not None:
            z = getattr(torch, self.final_act)(z)
        return z

    def forward(self, batch, n_id, msg, t, edge_index, id_mapper):
        P = self._P()
        z = self.embed(batch, n_id, msg, t, edge_index)
        src, dst = batch.src, batch.dst
        if self.tgb:                          # ctdg_models.py:533-538
            n_neg = batch.n_neg
            out = readout_ref(P, z[id_mapper[src]], z[id_mapper[dst]])
            return out[:-n_neg], out[-n_neg:], z[id_mapper[src[:-n_neg]]], z[id_mapper[dst[:-n_neg]]]
        if self.predict_dst:                  # ctdg_models.py:460-462
            h = F.linear(z[id_mapper[dst]], P["readout.lin.weight"], P["readout.lin.bias"]).relu()
            out = F.linear(h, P["readout.lin_final.weight"], P["readout.lin_final.bias"])
            return out, None, z[id_mapper[src]], z[id_mapper[dst]]
        neg_dst = getattr(batch, "neg_dst", None)   # ctdg_models.py:463-470
        pos_out = readout_ref(P, z[id_mapper[src]], z[id_mapper[dst]])
        neg_out = readout_ref(P, z[id_mapper[src]], z[id_mapper[neg_dst]]) if neg_dst is not None else None
        return pos_out, neg_out, z[id_mapper[src]], z[id_mapper[dst]]


class Batch:
    """Minimal attribute bag standing in for a PyG TemporalData mini-batch."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


# =============================================================================
# 5. Per-batch drivers restated from train_link.py (the caller contract)
# =============================================================================
def khop_lookup(loader, seeds: Tensor, num_hops: int):
    """train_link.py:258-262: n_id <- loader(n_id) K times, keep the last result."""
    n_id = seeds
    edge_index = torch.empty((2, 0), dtype=torch.long, device=seeds.device)
    e_id = torch.empty((0,), dtype=torch.long, device=seeds.device)
    for _ in range(num_hops):
        n_id, edge_index, e_id = loader(n_id)
    return n_id, edge_index, e_id


def train_step_ref(model: OracleCTAN, loader, helper: Tensor, optimizer, stream: dict, lo: int, hi: int,
                   neg_dst: Tensor, criterion=None):
    """One iteration of tgb_train (train_link.py:237-282) / train (train_link.py:297-343) on events
    [lo,hi) of `stream` (dict of src,dst,t,msg,x) with pre-drawn negatives.  Returns the loss tensor."""
    criterion = criterion or torch.nn.BCEWithLogitsLoss()
    src, dst, t, msg = stream["src"][lo:hi], stream["dst"][lo:hi], stream["t"][lo:hi], stream["msg"][lo:hi]
    optimizer.zero_grad()
    seeds = torch.cat([src, dst, neg_dst]).unique()
    n_id, edge_index, e_id = khop_lookup(loader, seeds, model.num_gnn_layers)
    helper[n_id] = torch.arange(n_id.numel(), device=n_id.device)
    if getattr(model, "tgb", type(model).__name__ == "CTAN_tgb"):      # also drives any model with the reference API
        batch = Batch(src=torch.cat([src, src]), dst=torch.cat([dst, neg_dst]), n_neg=neg_dst.numel(), x=stream.get("x"))
    else:
        batch = Batch(src=src, dst=dst, neg_dst=neg_dst, x=stream.get("x"))
    pos_out, neg_out, src_emb, dst_emb = model(batch, n_id, stream["msg"][e_id], stream["t"][e_id], edge_index, helper)
    loss = criterion(pos_out, torch.ones_like(pos_out)) + criterion(neg_out, torch.zeros_like(neg_out))
    model.update(src, dst, t, msg, src_emb, dst_emb)
    loader.insert(src, dst)
    loss.backward()
    optimizer.step()
    model.detach_memory()
    return loss.detach(), pos_out.detach(), neg_out.detach()


@torch.no_grad()
def infer_step_ref(model: OracleCTAN, loader, helper: Tensor, stream: dict, lo: int, hi: int, neg_dst: Tensor):
    """One iteration of eval (train_link.py:361-405) / tgb_test 'train' split (161-195)."""
    src, dst, t, msg = stream["src"][lo:hi], stream["dst"][lo:hi], stream["t"][lo:hi], stream["msg"][lo:hi]
    seeds = torch.cat([src, dst, neg_dst]).unique()
    n_id, edge_index, e_id = khop_lookup(loader, seeds, model.num_gnn_layers)
    helper[n_id] = torch.arange(n_id.numel(), device=n_id.device)
    if getattr(model, "tgb", type(model).__name__ == "CTAN_tgb"):      # also drives any model with the reference API
        batch = Batch(src=torch.cat([src, src]), dst=torch.cat([dst, neg_dst]), n_neg=neg_dst.numel(), x=stream.get("x"))
    else:
        batch = Batch(src=src, dst=dst, neg_dst=neg_dst, x=stream.get("x"))
    pos_out, neg_out, src_emb, dst_emb = model(batch, n_id, stream["msg"][e_id], stream["t"][e_id], edge_index, helper)
    model.update(src, dst, t, msg, src_emb, dst_emb)
    loader.insert(src, dst)
    return pos_out, neg_out


@torch.no_grad()
def tgb_eval_batch_ref(model: OracleCTAN, loader, helper: Tensor, stream: dict, lo: int, hi: int,
                       neg_lists, fast: bool = False):
    """One batch of tgb_test, split_mode 'val'/'test' (train_link.py:111-159) or 'val_fast'
    (60-109): one model forward PER QUERY, numpy MRR per query, then one memory update + insert
    with the concatenated per-query embeddings (197-199).  `neg_lists[q]` = negatives of query q.
    Returns (list of per-query MRR, pos scores, list of neg scores)."""
    assert model.tgb
    src, dst, t, msg = stream["src"][lo:hi], stream["dst"][lo:hi], stream["t"][lo:hi], stream["msg"][lo:hi]
    mrrs, pos_s, neg_s, s_embs, d_embs = [], [], [], [], []
    for q in range(hi - lo):
        neg = torch.as_tensor(neg_lists[q], dtype=torch.long, device=src.device)
        s_rep = torch.full((1 + neg.numel(),), int(src[q]), dtype=torch.long, device=src.device)
        if fast:
            seeds = torch.cat([src[q].view(1), dst[q].view(1), neg]).unique()
        else:
            seeds = torch.cat([s_rep, dst, neg]).unique()      # NB: the WHOLE batch's pos_dst (train_link.py:125)
        n_id, edge_index, e_id = khop_lookup(loader, seeds, model.num_gnn_layers)
        helper[n_id] = torch.arange(n_id.numel(), device=n_id.device)
        b = Batch(src=s_rep, dst=torch.cat([dst[q].view(1), neg]), n_neg=neg.numel(), x=stream.get("x"))
        pos_out, neg_out, s_emb, d_emb = model(b, n_id, stream["msg"][e_id], stream["t"][e_id], edge_index, helper)
        mrrs.append(mrr_ref(pos_out.squeeze(-1).cpu().numpy(), neg_out.squeeze(-1).cpu().numpy()))
        pos_s.append(pos_out.squeeze(-1)); neg_s.append(neg_out.squeeze(-1))
        s_embs.append(s_emb); d_embs.append(d_emb)
    model.update(src, dst, t, msg, torch.cat(s_embs, 0), torch.cat(d_embs, 0))
    loader.insert(src, dst)
    return mrrs, torch.cat(pos_s), neg_s


# ------------------------------------------------------------------ C1: sequence classification loops
def sequence_pass_ref(model, loader, helper: Tensor, seqs: list, x: Tensor, meta_batches: list, optimizer=None,
                      criterion=None, reset: bool = True):
    """train (optimizer given; train_sequence.py:16-65) or eval (optimizer None; train_sequence.py:68-110)
    over path-graph sequences.  `seqs[i]` = dict(src,dst,t,msg [L-1,...], y [1]); `x` [1,num_nodes,F] is shared.
    One event per step, K lookups with `e_id % (L-1)`, only the LAST prediction of a sequence enters the
    loss; memory / neighbour state is reset at the start of a training pass only.  Works for any
    (model, loader) pair with the reference API.  Returns (logits [n_seq], labels [n_seq])."""
    criterion = criterion or torch.nn.BCEWithLogitsLoss()
    dev = helper.device
    train = optimizer is not None
    if train and reset:
        model.reset_memory()
        loader.reset_state()
    all_pred, all_true = [], []
    with torch.set_grad_enabled(train):
        for idx in meta_batches:
            if train:
                optimizer.zero_grad()
            preds, trues = [], []
            for i in idx:
                d = seqs[int(i)]
                msg_all, t_all = d["msg"].to(dev), d["t"].to(dev)
                last = None
                for e in range(d["src"].numel()):
                    src, dst = d["src"][e:e + 1].to(dev), d["dst"][e:e + 1].to(dev)
                    n_id = torch.cat([src, dst]).unique()
                    edge_index = torch.empty((2, 0), dtype=torch.long, device=dev)
                    e_id = loader.e_id[n_id]
                    for _ in range(model.num_gnn_layers):
                        n_id, edge_index, e_id = loader(n_id)
                        e_id = e_id % msg_all.shape[0]
                    helper[n_id] = torch.arange(n_id.size(0), device=dev)
                    y_pred, _, se, de = model(Batch(src=src, dst=dst, x=x.to(dev)), n_id, msg_all[e_id], t_all[e_id],
                                              edge_index, helper)
                    last = y_pred.squeeze(0)
                    model.update(src, dst, t_all[e:e + 1], msg_all[e:e + 1], se, de)
                    loader.insert(src, dst)
                preds.append(last)
                trues.append(d["y"].to(dev))
            if train:
                loss = criterion(torch.cat(preds, -1), torch.cat(trues, -1))
                loss.backward()
                optimizer.step()
                model.detach_memory()
            all_pred += [p.detach() for p in preds]
            all_true += trues
    return torch.cat(all_pred), torch.cat(all_true)


def path_graph_events(seq_len: int, n_seq: int):
    """src, dst, t of the Temporal Path Graph generator (datasets/label_prediction_sequence.py:27-57):
    sequence i is the chain i*L -> i*L+1 -> ... with t = 0..L-2."""
    base = torch.arange(n_seq).view(-1, 1) * seq_len
    k = torch.arange(seq_len - 1).view(1, -1)
    return (base + k).reshape(-1), (base + k + 1).reshape(-1), k.expand(n_seq, -1).reshape(-1)


def delta_t_stats_ref(src, dst, t, init_time=0):
    """`compute_stats` (utils.py:46-92): mean / population std (float64) of `t - last_seen[node]` over BOTH
    endpoints of every event in order; both endpoints of an event see the table as it was before the event;
    a node's first occurrence is measured from init_time."""
    last = {}
    diffs = []
    for s, d, tt in zip(src.tolist(), dst.tolist(), t.tolist()):
        diffs.append(tt - last.get(s, init_time))
        diffs.append(tt - last.get(d, init_time))
        last[s] = tt
        last[d] = tt
    a = np.asarray(diffs, dtype=np.float64)
    return float(a.mean()), float(a.std())
