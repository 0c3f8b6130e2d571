"""PyG / torch-scatter shim -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Purpose: the reference (`/root/reference`) is pure Python but its operators come from
torch-geometric 2.3.1 and torch-scatter 2.1.2 (`env.yml:100-101`), neither of which is
installed in this image (no network).  This module registers minimal stand-ins in
``sys.modules`` so that the reference's OWN files

    models/ctdg_models.py, models/gnn_layers.py, models/memory_layers.py,
    models/predictors.py, TGB/modules/{neighbor_loader,time_enc,msg_agg}.py

import and run UNMODIFIED from where they lie.  It is used only
  * by ``tests/golden/make_golden.py`` (run in the build container, where /root/reference
    exists) to generate the committed golden vectors, and
  * by ``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent).

What is restated here (from the published PyG 2.3.1 algorithm; source not available
offline -- SURVEY.md App. A.4/A.5 lists the assumptions):
  TransformerConv (heads=1 path used by `models/gnn_layers.py:82`), AntiSymmetricConv
  (`models/gnn_layers.py:86`), utils.softmax, utils.scatter, inits.zeros/ones,
  resolver.activation_resolver, torch_scatter.scatter_max, data.TemporalData,
  loader.TemporalDataLoader.  TimeEncoder / LastAggregator / LastNeighborLoader are NOT
  restated: the reference vendors them verbatim under TGB/modules/ and we load those files.
"""
import importlib.util
import math
import os
import sys
import types

import torch
from torch import Tensor
from torch.nn import Parameter

REFERENCE_ROOT = os.environ.get("CTAN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


# ----------------------------------------------------------------------------- utils
def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size=None, reduce: str = "sum") -> Tensor:
    """torch_geometric.utils.scatter (2.3.1) for dim=0 and reduce in {sum, add, max, mean}."""
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = (dim_size,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    if reduce in ("sum", "add"):
        return src.new_zeros(size).scatter_add_(0, idx, src)
    if reduce == "mean":
        out = src.new_zeros(size).scatter_add_(0, idx, src)
        cnt = src.new_zeros(dim_size).scatter_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        return out / cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))
    if reduce == "max":
        return src.new_zeros(size).scatter_reduce_(0, idx, src, reduce="amax", include_self=False)
    raise ValueError(reduce)


def softmax(src: Tensor, index: Tensor, ptr=None, num_nodes=None, dim: int = 0) -> Tensor:
    """torch_geometric.utils.softmax (2.3.1): segment softmax with +1e-16 in the denominator."""
    n = num_nodes if num_nodes is not None else (int(index.max()) + 1 if index.numel() else 0)
    src_max = scatter(src.detach(), index, 0, n, reduce="max")
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = scatter(out, index, 0, n, reduce="sum") + 1e-16
    return out / out_sum.index_select(0, index)


def scatter_max(src: Tensor, index: Tensor, dim: int = 0, dim_size=None):
    """torch_scatter.scatter_max CPU semantics: strict '>' update => first arg-max wins;
    empty groups get arg = src.size(dim)."""
    assert dim == 0 and src.dim() == 1
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    out = src.new_zeros(dim_size)
    arg = torch.full((dim_size,), src.size(0), dtype=torch.long, device=src.device)
    if src.numel() == 0:
        return out, arg
    out = out.scatter_reduce(0, index, src, reduce="amax", include_self=False)
    is_max = src == out.index_select(0, index)
    pos = torch.arange(src.size(0), device=src.device)
    cand = torch.where(is_max, pos, torch.full_like(pos, src.size(0)))
    arg = arg.scatter_reduce(0, index, cand, reduce="amin", include_self=True)
    return out, arg


def zeros(t):
    if t is not None:
        t.data.fill_(0.0)


def ones(t):
    if t is not None:
        t.data.fill_(1.0)


def activation_resolver(query="relu", *args, **kwargs):
    if query is None or (callable(query) and not isinstance(query, str)):
        return query
    table = {"tanh": torch.nn.Tanh, "relu": torch.nn.ReLU, "sigmoid": torch.nn.Sigmoid,
             "elu": torch.nn.ELU, "leakyrelu": torch.nn.LeakyReLU}
    return table[query.lower().replace("_", "")](*args, **kwargs)


# ----------------------------------------------------------------------------- nn
class PygLinear(torch.nn.Module):
    """torch_geometric.nn.dense.linear.Linear with default initialisers:
    weight ~ kaiming_uniform(fan=in, a=sqrt(5)), bias ~ U(+-1/sqrt(in))."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        bound = math.sqrt(6 / ((1 + math.sqrt(5) ** 2) * self.in_channels))
        self.weight.data.uniform_(-bound, bound)
        if self.bias is not None:
            b = 1.0 / math.sqrt(self.in_channels)
            self.bias.data.uniform_(-b, b)

    def forward(self, x):
        return torch.nn.functional.linear(x, self.weight, self.bias)


class TransformerConv(torch.nn.Module):
    """PyG 2.3.1 TransformerConv, restricted to what CTAN uses (heads=1, concat, beta=False,
    dropout=0, aggr='add'); `root_weight=False` drops the skip term but `lin_skip` still exists."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, beta=False, dropout=0.0,
                 edge_dim=None, bias=True, root_weight=True, **kwargs):
        super().__init__()
        assert heads == 1 and concat and not beta and dropout == 0.0
        assert kwargs.get("aggr", "add") in ("add", "sum", "max")
        self.aggr = kwargs.get("aggr", "add")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.root_weight, self.edge_dim = root_weight, edge_dim
        self.lin_key = PygLinear(in_channels, heads * out_channels)
        self.lin_query = PygLinear(in_channels, heads * out_channels)
        self.lin_value = PygLinear(in_channels, heads * out_channels)
        self.lin_edge = PygLinear(edge_dim, heads * out_channels, bias=False) if edge_dim is not None else None
        self.lin_skip = PygLinear(in_channels, heads * out_channels, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        self.lin_key.reset_parameters()
        self.lin_query.reset_parameters()
        self.lin_value.reset_parameters()
        if self.lin_edge is not None:
            self.lin_edge.reset_parameters()
        self.lin_skip.reset_parameters()

    def forward(self, x, edge_index, edge_attr=None):
        C = self.out_channels
        query, key, value = self.lin_query(x), self.lin_key(x), self.lin_value(x)
        j, i = edge_index[0], edge_index[1]            # source j -> target i
        key_j, value_j, query_i = key[j], value[j], query[i]
        if self.lin_edge is not None:
            e = self.lin_edge(edge_attr)
            key_j = key_j + e
        alpha = (query_i * key_j).sum(dim=-1) / math.sqrt(C)
        alpha = softmax(alpha, i, num_nodes=x.size(0))
        out = value_j
        if self.lin_edge is not None:
            out = out + e
        out = out * alpha.view(-1, 1)
        out = scatter(out, i, 0, x.size(0), reduce="sum" if self.aggr != "max" else "max")
        if self.root_weight:
            out = out + self.lin_skip(x)
        return out


class AntiSymmetricConv(torch.nn.Module):
    """PyG 2.3.1 AntiSymmetricConv: x <- x + eps * act(x (W - W^T - gamma I)^T + phi(x) + b)."""

    def __init__(self, in_channels, phi=None, num_iters=1, epsilon=0.1, gamma=0.1, act="tanh",
                 act_kwargs=None, bias=True):
        super().__init__()
        assert phi is not None
        self.in_channels, self.num_iters, self.gamma, self.epsilon = in_channels, num_iters, gamma, epsilon
        self.act = activation_resolver(act, **(act_kwargs or {}))
        self.W = Parameter(torch.empty(in_channels, in_channels))
        self.register_buffer("eye", torch.eye(in_channels))
        self.phi = phi
        self.bias = Parameter(torch.empty(in_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.kaiming_uniform_(self.W, a=math.sqrt(5))
        self.phi.reset_parameters()
        zeros(self.bias)

    def forward(self, x, edge_index, *args, **kwargs):
        antisymmetric_W = self.W - self.W.t() - self.gamma * self.eye
        for _ in range(self.num_iters):
            h = self.phi(x, edge_index, *args, **kwargs)
            h = x @ antisymmetric_W.t() + h
            if self.bias is not None:
                h += self.bias
            if self.act is not None:
                h = self.act(h)
            x = x + self.epsilon * h
        return x


class IdentityMessage(torch.nn.Module):
    def __init__(self, raw_msg_dim, memory_dim, time_dim):
        super().__init__()
        self.out_channels = raw_msg_dim + 2 * memory_dim + time_dim


class TGNMemory(torch.nn.Module):  # base-class stub only (TGN/DyRep memories are out of scope)
    def __init__(self, *a, **k):
        super().__init__()


# ----------------------------------------------------------------------------- data
class TemporalData:
    """Attribute bag with PyG TemporalData slicing semantics."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_events(self):
        return self.src.size(0)

    @property
    def num_nodes(self):
        return int(max(int(self.src.max()), int(self.dst.max()))) + 1

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def __contains__(self, k):
        return k in self.__dict__

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return getattr(self, idx)
        n = self.num_events
        out = TemporalData()
        for k in self.keys():
            v = getattr(self, k)
            if isinstance(v, Tensor) and v.dim() > 0 and v.size(0) == n:
                v = v[idx]
            setattr(out, k, v)
        return out

    def to(self, device):
        for k in self.keys():
            v = getattr(self, k)
            if isinstance(v, Tensor):
                setattr(self, k, v.to(device))
        return self

    def train_val_test_split(self, val_ratio: float = 0.15, test_ratio: float = 0.15):
        """PyG 2.3.1: chronological cut at the (1-val-test) and (1-test) quantiles of t (counts of t <= quantile)."""
        import numpy as np
        val_time, test_time = np.quantile(self.t.cpu().numpy(), [1.0 - val_ratio - test_ratio, 1.0 - test_ratio])
        val_idx = int((self.t <= val_time).sum())
        test_idx = int((self.t <= test_time).sum())
        return self[:val_idx], self[val_idx:test_idx], self[test_idx:]


class TemporalDataLoader:
    def __init__(self, data, batch_size=1, **kwargs):
        self.data, self.batch_size = data, batch_size

    def __len__(self):
        return (self.data.num_events + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for i in range(0, self.data.num_events, self.batch_size):
            yield self.data[i:i + self.batch_size]


# ----------------------------------------------------------------------------- install
def _load_file(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_installed = None


class stable_sort:
    """Context manager: make `Tensor.sort()` stable while the reference runs.

    `LastNeighborLoader.insert` (TGB/modules/neighbor_loader.py:55) calls `nodes.sort()` without
    `stable=True`; torch's CPU sort is only stable for tiny inputs, so when one node receives more
    than `size` events in one batch the surviving entries are IMPLEMENTATION-DEFINED in the
    reference (and racy on CUDA).  The contract adopted by this project is the stable-sort
    behaviour, which the reference exhibits whenever its sort happens to be
    stable; this patch pins the reference to it when generating goldens.  Nothing else changes."""

    def __enter__(self):
        self._orig = torch.Tensor.sort

        def _sort(t, *args, **kwargs):
            if not args and "stable" not in kwargs:
                kwargs["stable"] = True
            elif args and not isinstance(args[0], bool) and "stable" not in kwargs:
                kwargs["stable"] = True
            return self._orig(t, *args, **kwargs)

        torch.Tensor.sort = _sort
        return self

    def __exit__(self, *exc):
        torch.Tensor.sort = self._orig
        return False


def install():
    """Register the stand-ins and import the reference's model package.  Returns a namespace
    with the reference's own classes (CTAN, CTAN_tgb, SimpleMemory, LastNeighborLoader, ...)."""
    global _installed
    if _installed is not None:
        return _installed
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    def mk(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mk("torch_scatter", scatter_max=scatter_max)
    tg = mk("torch_geometric")
    tg.utils = mk("torch_geometric.utils", scatter=scatter, softmax=softmax)
    tg.data = mk("torch_geometric.data", TemporalData=TemporalData)
    tg.loader = mk("torch_geometric.loader", TemporalDataLoader=TemporalDataLoader)
    inits = mk("torch_geometric.nn.inits", zeros=zeros, ones=ones)
    resolver = mk("torch_geometric.nn.resolver", activation_resolver=activation_resolver)

    # vendored verbatim copies of the PyG TGN helpers (TGB/modules/*.py)
    tgbm = os.path.join(REFERENCE_ROOT, "TGB", "modules")
    nl = _load_file("_ref_neighbor_loader", os.path.join(tgbm, "neighbor_loader.py"))
    te = _load_file("_ref_time_enc", os.path.join(tgbm, "time_enc.py"))
    ma = _load_file("_ref_msg_agg", os.path.join(tgbm, "msg_agg.py"))

    tgn = mk("torch_geometric.nn.models.tgn", IdentityMessage=IdentityMessage,
             LastAggregator=ma.LastAggregator, TimeEncoder=te.TimeEncoder,
             LastNeighborLoader=nl.LastNeighborLoader)
    models_pkg = mk("torch_geometric.nn.models", tgn=tgn)
    tg.nn = mk("torch_geometric.nn", AntiSymmetricConv=AntiSymmetricConv, TransformerConv=TransformerConv,
               TGNMemory=TGNMemory, inits=inits, resolver=resolver, models=models_pkg)

    # the reference's own `models` package, loaded from where it lies
    spec = importlib.util.spec_from_file_location(
        "ref_models", os.path.join(REFERENCE_ROOT, "models", "__init__.py"),
        submodule_search_locations=[os.path.join(REFERENCE_ROOT, "models")])
    ref_models = importlib.util.module_from_spec(spec)
    sys.modules["ref_models"] = ref_models
    spec.loader.exec_module(ref_models)
    cm = sys.modules["ref_models.ctdg_models"]

    # TGB evaluator's MRR (numpy branch) lives in a file with heavy imports; load only if importable
    ns = types.SimpleNamespace(
        CTAN=cm.CTAN, CTAN_tgb=cm.CTAN_tgb, SimpleMemory=cm.SimpleMemory, CTANEmbedding=cm.CTANEmbedding,
        LinkPredictor=cm.LinkPredictor, TGBLinkPredictor=cm.TGBLinkPredictor,
        SequencePredictor=cm.SequencePredictor, LastNeighborLoader=nl.LastNeighborLoader,
        TimeEncoder=te.TimeEncoder, LastAggregator=ma.LastAggregator, TemporalData=TemporalData,
        TemporalDataLoader=TemporalDataLoader, models=ref_models)
    _installed = ns
    return ns
