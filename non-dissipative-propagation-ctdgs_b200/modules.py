"""Host-side mirror of the reference's module API for the CTAN path, running on our kernels.

Same class names, constructor arguments, forward contracts and state-dict keys as the reference
(SURVEY.md S8b, App. A.8):
    CTAN, CTAN_tgb                 models/ctdg_models.py:382-470, 474-540
    CTANEmbedding                  models/gnn_layers.py:59-98
    SimpleMemory                   models/memory_layers.py:117-156
    LinkPredictor, TGBLinkPredictor, SequencePredictor   models/predictors.py:4-48
Parameters are initialised with the same draws in the same order as the reference (App. A.7), so a
model built under the same torch seed has identical weights.  There is no CPU path."""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, Optional, Union

import torch

from . import _lib
from ._lib import call, ptr
from .functional import ctan_embed, link_readout
from .neighbor_loader import cached_rowptr


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.CtanLibraryError(f"{what}: ctan_b200 runs on CUDA only (got a {t.device} tensor); "
                                    "call model.to('cuda') -- there is no CPU fallback")


class _PygLinear(torch.nn.Module):
    """Parameter holder with torch_geometric.nn.dense.Linear's default initialisation."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        bound = math.sqrt(6 / ((1 + math.sqrt(5) ** 2) * self.in_channels))
        self.weight.data.uniform_(-bound, bound)
        if self.bias is not None:
            b = 1.0 / math.sqrt(self.in_channels)
            self.bias.data.uniform_(-b, b)


class _TransformerConvParams(torch.nn.Module):
    """The five Linears of PyG's TransformerConv(heads=1, root_weight=False) -- `lin_skip` exists but is
    unused, exactly as upstream.  The arithmetic lives in csrc/edge.cu + csrc/dense.cu."""

    def __init__(self, channels: int, edge_dim: int):
        super().__init__()
        self.lin_key = _PygLinear(channels, channels)
        self.lin_query = _PygLinear(channels, channels)
        self.lin_value = _PygLinear(channels, channels)
        self.lin_edge = _PygLinear(edge_dim, channels, bias=False)
        self.lin_skip = _PygLinear(channels, channels)
        self.reset_parameters()

    def reset_parameters(self):
        for m in (self.lin_key, self.lin_query, self.lin_value, self.lin_edge, self.lin_skip):
            m.reset_parameters()


class _AntiSymmetricParams(torch.nn.Module):
    def __init__(self, channels: int, phi: _TransformerConvParams, num_iters: int, epsilon: float, gamma: float):
        super().__init__()
        self.in_channels, self.num_iters, self.epsilon, self.gamma = channels, num_iters, epsilon, gamma
        self.W = torch.nn.Parameter(torch.empty(channels, channels))
        self.register_buffer("eye", torch.eye(channels))
        self.phi = phi
        self.bias = torch.nn.Parameter(torch.empty(channels))
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.kaiming_uniform_(self.W, a=math.sqrt(5))
        self.phi.reset_parameters()
        self.bias.data.zero_()


class _TimeEncoderParams(torch.nn.Module):
    def __init__(self, out_channels: int):
        super().__init__()
        self.out_channels = out_channels
        self.lin = torch.nn.Linear(1, out_channels)

    def reset_parameters(self):
        self.lin.reset_parameters()


class CTANEmbedding(torch.nn.Module):
    def __init__(self, in_channels: int, node_dim: int, msg_dim: int, time_dim: int, num_iters: int = 1,
                 act: Union[str, Callable, None] = "tanh", act_kwargs: Optional[Dict[str, Any]] = None,
                 epsilon: float = 0.1, gamma: float = 0.1, mean_delta_t: float = 0., std_delta_t: float = 1.,
                 conv_type: str = "TransformerConv"):
        assert conv_type in ["TransformerConv"]
        if act != "tanh" or act_kwargs:
            raise NotImplementedError("ctan_b200 implements gnn_act='tanh' (the only value in the reference confs)")
        super().__init__()
        self.mean_delta_t, self.std_delta_t = mean_delta_t, std_delta_t
        self.out_channels = in_channels
        self.enc_x = torch.nn.Linear(in_channels + node_dim, in_channels)
        phi = _TransformerConvParams(in_channels, msg_dim + time_dim)
        self.aconv = _AntiSymmetricParams(in_channels, phi, num_iters, epsilon, gamma)
        self.time_enc = _TimeEncoderParams(time_dim)

    def reset_parameters(self):
        self.time_enc.reset_parameters()
        self.aconv.reset_parameters()

    def forward(self, x, last_update, edge_index, t, msg, *, n_id=None, rowptr=None, final_tanh=False):
        """Reference contract (gnn_layers.py:93): x [N, D+Fn], last_update [N] (local), edge_index [2,E]
        destination-sorted, t [E], msg [E,Fe].  Keyword extras: `n_id` makes `last_update` the global
        table; `rowptr` skips the CSR rebuild; `final_tanh` fuses CTAN's final activation."""
        _require_cuda(x, "CTANEmbedding.forward")
        N = x.size(0)
        if rowptr is None:
            rowptr = cached_rowptr(edge_index, N)
        if rowptr is None:
            rowptr = _rowptr_from_edge_index(edge_index, N)
        a, p = self.aconv, self.aconv.phi
        return ctan_embed(x, last_update, n_id, edge_index[0], rowptr, msg, t, num_iters=a.num_iters,
                          epsilon=a.epsilon, gamma=a.gamma, mean_delta_t=self.mean_delta_t,
                          std_delta_t=self.std_delta_t, final_tanh=final_tanh, Wenc=self.enc_x.weight,
                          benc=self.enc_x.bias, Wq=p.lin_query.weight, bq=p.lin_query.bias, Wk=p.lin_key.weight,
                          bk=p.lin_key.bias, Wv=p.lin_value.weight, bv=p.lin_value.bias, We=p.lin_edge.weight,
                          W=a.W, b=a.bias, tw=self.time_enc.lin.weight, tb=self.time_enc.lin.bias)


def _rowptr_from_edge_index(edge_index: torch.Tensor, N: int) -> torch.Tensor:
    """CSR pointer for an edge list that did not come from our loader; it must be sorted by destination
    (the layout every reference call site provides).  One host sync to validate."""
    dst = edge_index[1].contiguous()
    E = dst.numel()
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=dst.device)
    flag = torch.zeros(1, dtype=torch.int32, device=dst.device)
    with torch.cuda.device(dst.device):
        call("ctan_csr_from_sorted", ptr(dst, torch.long), E, N, ptr(rowptr), ptr(flag))
    if int(flag.item()) != 0:
        raise ValueError("edge_index[1] must be sorted ascending with ids in [0, N) "
                         "(the order LastNeighborLoader produces)")
    return rowptr


class SimpleMemory(torch.nn.Module):
    def __init__(self, num_nodes: int, memory_dim: int, aggregator_module: Optional[Callable] = None,
                 init_time: int = 0) -> None:
        super().__init__()
        self.num_nodes, self.memory_dim = num_nodes, memory_dim
        self.init_time = int(init_time)
        self.aggr_module = aggregator_module      # kept for signature parity; "last" aggregation is built in
        self.register_buffer("memory", torch.zeros(num_nodes, memory_dim))
        self.register_buffer("last_update", torch.ones(num_nodes, dtype=torch.long) * self.init_time)
        self.register_buffer("_assoc", torch.zeros(num_nodes, dtype=torch.long))

    def update_state(self, src, pos_dst, t, src_emb, pos_dst_emb):
        _require_cuda(self.memory, "SimpleMemory.update_state")
        B = src.numel()
        if B == 0:
            return
        se = src_emb.detach().float().contiguous()
        de = pos_dst_emb.detach().float().contiguous()
        sc, dc, tc = src.contiguous(), pos_dst.contiguous(), t.contiguous()     # alive across the launch
        with torch.cuda.device(self.memory.device):
            call("ctan_mem_update", ptr(self.memory), ptr(self.last_update), self.memory_dim,
                 ptr(sc, torch.long), ptr(dc, torch.long), ptr(tc, torch.long), B, ptr(se), ptr(de), 0, None, None)

    def reset_state(self):
        _require_cuda(self.memory, "SimpleMemory.reset_state")
        with torch.cuda.device(self.memory.device):
            call("ctan_mem_reset", ptr(self.memory), ptr(self.last_update), self.num_nodes, self.memory_dim,
                 self.init_time)

    def detach(self):
        self.memory.detach_()

    def forward(self, n_id):
        return self.memory[n_id], self.last_update[n_id]


class _Predictor(torch.nn.Module):
    def _run(self, z_src, z_dst):
        if z_src is None:
            z = z_dst
            P = z.size(0)
            idx = torch.arange(P, device=z.device)
            return link_readout(z, None, idx, None, None, None, self.lin.weight, self.lin.bias,
                                self.lin_final.weight, self.lin_final.bias)
        P = z_src.size(0)
        z = torch.cat([z_src, z_dst], 0)
        idx = torch.arange(2 * P, device=z.device)
        return link_readout(z, idx[:P], idx[P:], None, self.lin_src.weight, self.lin_src.bias,
                            self.lin_dst.weight, self.lin_dst.bias, self.lin_final.weight, self.lin_final.bias)


class TGBLinkPredictor(_Predictor):
    def __init__(self, in_channels):
        super().__init__()
        self.lin_src = torch.nn.Linear(in_channels, in_channels)
        self.lin_dst = torch.nn.Linear(in_channels, in_channels)
        self.lin_final = torch.nn.Linear(in_channels, 1)

    def forward(self, z_src, z_dst):
        return self._run(z_src, z_dst)


class LinkPredictor(_Predictor):
    def __init__(self, in_channels: int, hidden_channels: Optional[int] = None, out_channels: int = 1):
        super().__init__()
        if hidden_channels is None:
            hidden_channels = in_channels
        self.lin_src = torch.nn.Linear(in_channels, hidden_channels)
        self.lin_dst = torch.nn.Linear(in_channels, hidden_channels)
        self.lin_final = torch.nn.Linear(hidden_channels, out_channels)

    def forward(self, z_src, z_dst):
        return self._run(z_src, z_dst)


class SequencePredictor(_Predictor):
    def __init__(self, in_channels: int, hidden_channels: Optional[int] = None, out_channels: int = 1):
        super().__init__()
        if hidden_channels is None:
            hidden_channels = in_channels
        self.lin = torch.nn.Linear(in_channels, hidden_channels)
        self.lin_final = torch.nn.Linear(hidden_channels, out_channels)

    def forward(self, z_dst):
        return self._run(None, z_dst)


class CTAN(torch.nn.Module):
    def __init__(self, num_nodes: int, edge_dim: int, memory_dim: int, time_dim: int, node_dim: int = 0,
                 num_iters: int = 1, gnn_act: Union[str, Callable, None] = "tanh",
                 gnn_act_kwargs: Optional[Dict[str, Any]] = None, epsilon: float = 0.1, gamma: float = 0.1,
                 readout_hidden: Optional[int] = None, readout_out: Optional[int] = 1, mean_delta_t: float = 0.,
                 std_delta_t: float = 1., init_time: int = 0, final_act: Optional[str] = None,
                 predict_dst: bool = False, conv_type: str = "TransformerConv"):
        assert conv_type in ["TransformerConv"]
        if final_act not in (None, "tanh"):
            raise NotImplementedError("ctan_b200 implements final_act in {None, 'tanh'} (the reference confs)")
        super().__init__()
        self.memory = SimpleMemory(num_nodes, memory_dim, aggregator_module=None, init_time=init_time)
        self.gnn = CTANEmbedding(memory_dim, node_dim, edge_dim, time_dim, num_iters, gnn_act, gnn_act_kwargs,
                                 epsilon, gamma, mean_delta_t, std_delta_t, conv_type=conv_type)
        self.readout = (LinkPredictor(memory_dim, readout_hidden, readout_out) if not predict_dst
                        else SequencePredictor(memory_dim, readout_hidden, readout_out))
        self.gnn_act = gnn_act
        self.num_gnn_layers = num_iters
        self.num_nodes = num_nodes
        self.node_dim = node_dim
        self.predict_dst = predict_dst
        self.final_act = final_act
        self.layernorm = None

    # GenericModel plumbing (ctdg_models.py:24-34, 435-436)
    def reset_memory(self):
        self.memory.reset_state()

    def detach_memory(self):
        self.memory.detach()

    def update(self, src, pos_dst, t, msg, src_emb, pos_dst_emb):
        self.memory.update_state(src, pos_dst, t, src_emb, pos_dst_emb)

    def _embed(self, batch, n_id, msg, t, edge_index):
        mem = self.memory
        _require_cuda(mem.memory, "CTAN.forward")
        n_id = n_id.contiguous()
        N = n_id.numel()
        D = mem.memory_dim
        x = None
        if hasattr(batch, "x") and batch.x is not None:
            if batch.x.dim() == 3:
                x = batch.x.squeeze(0)
            elif batch.x.dim() == 2:
                x = batch.x
            else:
                raise ValueError(f"Unexpected node feature shape. Got {batch.x.shape}")
            x = x.to(device=mem.memory.device, dtype=torch.float32).contiguous()
        Fn = x.size(1) if x is not None else 0
        z0 = torch.empty((N, D + Fn), device=mem.memory.device)
        with torch.cuda.device(mem.memory.device):
            call("ctan_gather_rows2", ptr(mem.memory), D, ptr(x) if x is not None else None, Fn,
                 ptr(n_id, torch.long), N, None, ptr(z0))
        msg = msg.to(mem.memory.device)
        t = t.to(mem.memory.device)
        return self.gnn(z0, mem.last_update, edge_index, t, msg, n_id=n_id, final_tanh=self.final_act == "tanh")

    def forward(self, batch, n_id, msg, t, edge_index, id_mapper):
        src, pos_dst = batch.src, batch.dst
        neg_dst = batch.neg_dst if hasattr(batch, "neg_dst") else None
        z = self._embed(batch, n_id, msg, t, edge_index)
        r = self.readout
        if self.predict_dst:
            pos_out = link_readout(z, None, pos_dst, id_mapper, None, None, r.lin.weight, r.lin.bias,
                                   r.lin_final.weight, r.lin_final.bias)
            neg_out = None
        else:
            B = src.numel()
            if neg_dst is not None:
                out = link_readout(z, torch.cat([src, src]), torch.cat([pos_dst, neg_dst]), id_mapper,
                                   r.lin_src.weight, r.lin_src.bias, r.lin_dst.weight, r.lin_dst.bias,
                                   r.lin_final.weight, r.lin_final.bias)
                pos_out, neg_out = out[:B], out[B:]
            else:
                pos_out = link_readout(z, src, pos_dst, id_mapper, r.lin_src.weight, r.lin_src.bias,
                                       r.lin_dst.weight, r.lin_dst.bias, r.lin_final.weight, r.lin_final.bias)
                neg_out = None
        emb_src = z[id_mapper[src]]
        emb_pos_dst = z[id_mapper[pos_dst]]
        return pos_out, neg_out, emb_src, emb_pos_dst


class CTAN_tgb(CTAN):
    """CTAN with the TGB readout and the (pos ++ neg) batch convention (ctdg_models.py:474-540).
    Accepts the `separable` kwarg the TGB conf passes and defines `layernorm=None`."""

    def __init__(self, *args, separable: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        self.readout = TGBLinkPredictor(self.memory.memory_dim)

    def forward(self, batch, n_id, msg, t, edge_index, id_mapper):
        src, dst = batch.src, batch.dst
        n_neg = batch.n_neg
        z = self._embed(batch, n_id, msg, t, edge_index)
        r = self.readout
        if self.predict_dst:
            assert n_neg is None
            pos_out = link_readout(z, None, dst, id_mapper, None, None, r.lin.weight, r.lin.bias,
                                   r.lin_final.weight, r.lin_final.bias)
            return pos_out, None, z[id_mapper[src]], z[id_mapper[dst]]
        out = link_readout(z, src, dst, id_mapper, r.lin_src.weight, r.lin_src.bias, r.lin_dst.weight,
                           r.lin_dst.bias, r.lin_final.weight, r.lin_final.bias)
        pos_out, neg_out = out[:-n_neg], out[-n_neg:]
        emb_src = z[id_mapper[src[:-n_neg]]]
        emb_pos_dst = z[id_mapper[dst[:-n_neg]]]
        return pos_out, neg_out, emb_src, emb_pos_dst
