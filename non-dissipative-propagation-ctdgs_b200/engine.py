"""Zero-host-sync per-batch engine: the whole step of train_link.py:237-282 (tgb_train / train) or of
the no-grad loops (eval, tgb_test 'train' split: train_link.py:161-199, 361-405) as ONE stream of our
kernels -- lookup -> forward -> loss -> memory write-back + insert -> backward -> Adam -- with every
data-dependent size kept on the device, captured once in a CUDA graph and replayed per batch.

It drives the SAME module objects as the drop-in API (`CTAN`/`CTAN_tgb`, `LastNeighborLoader`): their
parameters are re-pointed into one flat buffer (so Adam is a single launch) and their state tables are
updated in place, so a run can switch between the engine and the reference-style Python loop.

Event stream (device resident): src, dst, t [n] int64, msg [n,Fe] fp32, neg [n, n_neg] int64 (negatives
are an INPUT, drawn by the caller exactly like train_link.py:245-251 / negative_sampler.py do), x
[num_nodes,Fn] fp32 or None.  Event ids of the loader are stream indices (cur_e_id counts from the
first event, train_link.py:265), so `msg[e_id]`/`t[e_id]` are gathers from the same tables."""
from __future__ import annotations

import contextlib
import os

import torch

from . import _lib
from ._lib import call, ptr
from .modules import CTAN, CTAN_tgb
from .neighbor_loader import LastNeighborLoader


def _param_list(model):
    """The 19 parameter slots the step uses, in the engine's fixed order (None where the head has no such tensor:
    SequencePredictor has a single input projection `lin`, models/predictors.py:37-48)."""
    g, a, r = model.gnn, model.gnn.aconv, model.readout
    p = a.phi
    if model.predict_dst:
        head = [None, None, r.lin.weight, r.lin.bias, r.lin_final.weight, r.lin_final.bias]
    else:
        head = [r.lin_src.weight, r.lin_src.bias, r.lin_dst.weight, r.lin_dst.bias, r.lin_final.weight, r.lin_final.bias]
    return [g.enc_x.weight, g.enc_x.bias, p.lin_query.weight, p.lin_query.bias, p.lin_key.weight, p.lin_key.bias,
            p.lin_value.weight, p.lin_value.bias, p.lin_edge.weight, a.W, a.bias, g.time_enc.lin.weight,
            g.time_enc.lin.bias] + head


LOSS_MODES = {"CrossEntropyLoss": 1, "MSELoss": 2, "L1Loss": 3, "BCEWithLogitsLoss": 4}


class StepEngine:
    def __init__(self, model: CTAN, loader: LastNeighborLoader, stream: dict, batch_size: int, n_neg: int = 1,
                 lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 use_graph: bool = True, hist_cap: int = 1 << 16, host_staged: bool = False, criterion=None,
                 out_cap: int = 0, neg_base: int = 0, y_base: int = 0, seq_mod: int = 0):
        """Heads (chosen from the model and `n_neg`):
          * link head with negatives (n_neg >= 1): pairs (src,dst) + (src,neg), fused BCE on pos/neg
            (train_link.py:269-270, 335-336) -- `criterion` must be None or BCEWithLogitsLoss(mean);
          * labelled head (n_neg == 0): pairs (src,dst) scored by LinkPredictor (readout_out >= 1) or, with
            predict_dst, z[dst] by SequencePredictor; loss = criterion(out, y) with y = stream["y"] rows of the batch:
            CrossEntropyLoss (int64 class ids; multiclass, train_link.py:708-709), MSELoss / L1Loss (regression,
            :706), BCEWithLogitsLoss (fp32 targets; train_sequence.py:57).
        out_cap > 0 keeps the scores of the last out_cap steps (`outputs(n)`); neg_base / y_base: event index of row 0
        of stream["neg"] / stream["y"] when those tables cover only a suffix of the event stream.
        seq_mod > 0: sequence mode (train_sequence.py): the stream is a pass over sequences of seq_mod events each; an
        edge's message / time row is (first row of the current sequence) + e_id % seq_mod (train_sequence.py:41); steps
        are issued one by one with `step(kind)`: "infer" | "grad" (accumulate gradients, scaled by set_grad_scale) |
        `adam()` (optimizer step on the accumulated gradients, then clear them)."""
        dev = model.memory.memory.device
        if dev.type != "cuda":
            raise _lib.CtanLibraryError("StepEngine needs the model on a CUDA device")
        _lib.lib()
        self.model, self.loader, self.dev = model, loader, dev
        self.tgb = isinstance(model, CTAN_tgb)
        self.B, self.n_neg = int(batch_size), int(n_neg)
        self.predict_dst = bool(model.predict_dst)
        self.labelled = self.n_neg == 0
        if self.predict_dst and not self.labelled:
            raise ValueError("predict_dst heads take no negatives (n_neg must be 0)")
        self.lr, self.betas = float(lr), (float(betas[0]), float(betas[1]))
        self.eps, self.wd = float(eps), float(weight_decay)
        self.use_graph, self.hist_cap, self.host_staged = use_graph, int(hist_cap), host_staged
        self.out_cap, self.neg_base, self.y_base = int(out_cap), int(neg_base), int(y_base)
        self.seq_mod = int(seq_mod)
        self.K = model.num_gnn_layers
        self.D = model.memory.memory_dim
        self.S = loader.size
        self.num_nodes = model.num_nodes
        self.final_tanh = model.final_act == "tanh"
        st = stream
        self.n_events = int(st["src"].numel())
        self.Fe = int(st["msg"].size(1))
        cname = type(criterion).__name__ if criterion is not None else None
        if self.labelled:
            if cname not in LOSS_MODES:
                raise NotImplementedError(f"labelled head: criterion must be one of {sorted(LOSS_MODES)}, got {cname}")
            if getattr(criterion, "reduction", "mean") != "mean" or getattr(criterion, "weight", None) is not None \
                    or getattr(criterion, "pos_weight", None) is not None or getattr(criterion, "label_smoothing", 0.0):
                raise NotImplementedError("labelled head: plain mean-reduced criteria only")
            self.loss_mode = LOSS_MODES[cname]
            if host_staged:
                raise NotImplementedError("host_staged mode covers the link head with negatives")
        else:
            if criterion is not None and (cname != "BCEWithLogitsLoss" or criterion.reduction != "mean"
                                          or criterion.pos_weight is not None or criterion.weight is not None):
                raise NotImplementedError("link head: criterion must be BCEWithLogitsLoss() (mean)")
            self.loss_mode = 0
        if host_staged:
            # batches arrive from the host: fixed device staging buffers + event tables filled as they come
            # two staging sets: the copy of batch j+1 (on a copy stream) overlaps the step of batch j
            self._stg = []
            for _ in range(2):                             # one packed buffer per set: [src | dst | t | neg]
                ids = torch.zeros((3 + self.n_neg) * self.B, dtype=torch.long, device=dev)
                Bq = self.B
                self._stg.append(dict(ids=ids, src=ids[:Bq], dst=ids[Bq:2 * Bq], t=ids[2 * Bq:3 * Bq],
                                      neg=ids[3 * Bq:].view(Bq, self.n_neg)))
            self.src, self.dst, self.t, self.neg = (self._stg[0][k] for k in ("src", "dst", "t", "neg"))
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._copied = [torch.cuda.Event(), torch.cuda.Event()]      # staging set filled
            self._consumed = [None, None]                                 # step that read the staging set is done
            self._consumed_evs = [torch.cuda.Event(), torch.cuda.Event()]
            self._push_par = self._run_par = 0
            self.t_tab = torch.zeros(self.n_events, dtype=torch.long, device=dev)
            self.msg = torch.zeros((self.n_events, self.Fe), device=dev)
            self._host_cursor = 0
        else:
            for k in ("src", "dst", "t", "msg") + (() if self.labelled else ("neg",)):
                if not st[k].is_cuda:
                    raise _lib.CtanLibraryError(f"stream['{k}'] must be device resident")
            self.src, self.dst, self.t = st["src"].contiguous(), st["dst"].contiguous(), st["t"].contiguous()
            self.t_tab = self.t
            self.msg = st["msg"].float().contiguous()
            if self.labelled:
                self.neg = None
            else:
                self.neg = st["neg"].contiguous().view(-1, self.n_neg)
        self.x = st.get("x")
        if self.x is not None:
            self.x = (self.x.squeeze(0) if self.x.dim() == 3 else self.x).to(dev).float().contiguous()
        self.Fn = self.x.size(1) if self.x is not None else 0
        self.T = model.gnn.time_enc.lin.bias.numel()
        r = model.readout
        self.Hd = (r.lin if self.predict_dst else r.lin_src).weight.size(0)
        self.O = r.lin_final.weight.size(0)
        if not self.labelled and self.O != 1:
            raise NotImplementedError("the link head with negatives is binary (readout_out=1)")
        self.y = None
        if self.labelled:
            y = st["y"]
            if not y.is_cuda:
                raise _lib.CtanLibraryError("stream['y'] must be device resident")
            if self.loss_mode == 1:
                self.y = y.reshape(-1).to(torch.long).contiguous()
            else:
                self.y = y.reshape(y.size(0), -1).float().contiguous()
                if self.y.size(1) != self.O:
                    raise ValueError(f"targets have {self.y.size(1)} columns, the head {self.O}")
        self.grad_scale = 1.0
        self.P = self.B * (1 + self.n_neg)
        self.n_seeds = self.B * (2 + self.n_neg)
        self.n_cap = min(self.num_nodes, self.n_seeds * (self.S + 1) ** self.K)
        c_cap = min(self.num_nodes, self.n_seeds * (self.S + 1) ** (self.K - 1))
        self.e_cap = max(c_cap * self.S, 1)
        # typical sizes (they size split-K grids and pick the batch-sized kernel variants; correctness never depends on
        # them): a quarter of the bound, but between 1x and 2x the seed count for nodes and 2x..4x for edges -- multi-hop
        # neighbourhoods overlap heavily (wiki, 600 seeds: N = 640 / 816 / 844 and E = 1.3 k / 2.8 k / 3.0 k at 1 / 3 / 5
        # hops against bounds of 3 600 / 9 227 nodes and 3 k / 46 k edges).  An OVER-estimate is what hurts: the K split
        # of the weight-gradient contractions becomes too coarse (lin_edge gradient at 3 hops: 60 us, 30 CTAs x 38 gather
        # rounds, with the old e_cap/4 guess).  retune() replaces the guesses by observed sizes.
        self.n_typ = min(self.n_cap, max(self.n_seeds, min(self.n_cap // 4, 2 * self.n_seeds)))
        self.e_typ = min(self.e_cap, max(2 * self.n_seeds, min(self.e_cap // 4, 4 * self.n_seeds)))
        self._flatten_params()
        self._alloc()
        self._graphs = {}
        # critical path on a high-priority stream, off-critical-path branches on normal-priority streams: when
        # both have CTAs pending, the block scheduler serves the critical path first
        self._main = torch.cuda.Stream(device=dev, priority=-1)
        self._side = [torch.cuda.Stream(device=dev, priority=0) for _ in range(4)]

    def _derive_choices(self):
        """Kernel variants and grid hints that follow from the typical sizes n_typ / e_typ."""
        D = self.D
        # a batch's edge set is small (~1e3 rows): the all-loads-in-flight variant of the edge projection (csrc/dense.cu)
        small = (self.e_typ <= 8192 and (self.Fe + self.T) % 4 == 0 and D in (32, 64, 128, 256)
                 and (D + 16) * (self.Fe + self.T + 4) * 4 <= 220 * 1024 and os.environ.get("CTAN_EPROJ_ROWS", "1") == "1")
        self._eproj_fn = "ctan_eproj_fwd_rows" if small else "ctan_eproj_fwd"
        # the binary link readout (hidden + head + loss + gradients wrt the embeddings) as one launch
        pt = 256 // self.Hd if self.Hd in (32, 64, 128, 256) else 0
        self._link_fused = (not self.labelled and self.O == 1 and pt > 0 and D % 4 == 0 and self.n_neg >= 1
                            and (2 * self.Hd * (D + 4) + pt * (2 * D + self.Hd) + 8) * 4 <= 200 * 1024
                            and os.environ.get("CTAN_LINK_FUSED", "1") == "1")
        # batch-sized backward side branches: W's gradient accumulated straight through A = W - W^T - gamma I (not in
        # gradient-accumulation mode, where dW is only formed at the end), time-encoder gradients straight from dEP
        self._wgrad_direct = (not self.seq_mod) and os.environ.get("CTAN_WGRAD_DIRECT", "1") == "1"
        self._tenc_fused = (self.e_typ <= 16384 and D % 32 == 0 and D <= 256 and (self.T * D + 16 * self.T) * 4 <= 96 * 1024
                            and os.environ.get("CTAN_TENC_FUSED", "1") == "1")
        # expected row count of the node-side contractions, packed into the `mode` argument of the tcgen05 entry points:
        # their CTAs need a whole SM each, so the persistent grid is sized by the typical N, not by the workspace bound
        self._row_hint = (self.n_typ << 8) if os.environ.get("CTAN_TC_ROW_HINT", "1") == "1" else 0

    def retune(self, n_nodes: int = 0, n_edges: int = 0):
        """Replace the typical-size guesses by observed sizes (default: the last batch's, read from the device -- one
        synchronisation) with 25 % head-room, re-derive the kernel choices and drop the captured graphs (they are
        re-captured on the next run).  Results never depend on the hints; call it once the stream is in its steady state
        if the subgraph sizes are far from 1-2x the seed count."""
        if n_nodes <= 0 or n_edges <= 0:
            n_nodes, n_edges = (int(v) for v in self.w["counts"][:2].tolist())
        self.n_typ = min(self.n_cap, max(64, n_nodes + n_nodes // 4))
        self.e_typ = min(self.e_cap, max(64, n_edges + n_edges // 4))
        self._derive_choices()
        self._graphs = {}

    # ------------------------------------------------------------------ parameters in one flat buffer
    def _flatten_params(self):
        self._slots = _param_list(self.model)
        plist = [p for p in self._slots if p is not None]
        self._plist = plist
        n = sum(p.numel() for p in plist)
        # ONE flat parameter buffer per model, shared by every engine built on it (a second engine must not re-home
        # the parameters under the first one's feet)
        shared = self.model.__dict__.get("_ctan_flat")
        reuse = False
        if shared is not None and shared.numel() == n and shared.device == self.dev:
            o, reuse = 0, True
            for p in plist:
                reuse = reuse and p.data.data_ptr() == shared.data_ptr() + 4 * o
                o += p.numel()
        self.flat_p = shared if reuse else torch.empty(n, device=self.dev)
        self.model.__dict__["_ctan_flat"] = self.flat_p
        # everything the backward accumulates into lives in ONE buffer so a single memset clears it
        N, D = self.n_cap, self.D
        K = self.K

        def pad(v):                      # every region starts on a 256-byte boundary (128-bit loads, TMA bases)
            return (v + 63) // 64 * 64
        o_dA = pad(n)
        o_dZ = o_dA + pad(D * D)
        self._o_ws = o_dZ                                      # [0, _o_ws): parameter gradients + dA; beyond: per-step scratch
        self._accumulate = False
        o_DQ = o_dZ + pad(N * D)
        # one DQKVA buffer PER anti-symmetric iteration: the backward of iteration k-1 does not wait for iteration k's
        # weight-gradient contraction (a side branch that reads DQKVA) nor re-clear the buffer on the critical path
        self.nDQ = K if os.environ.get("CTAN_DQ_MULTI", "1") == "1" else 1
        o_DX = o_DQ + self.nDQ * pad(N * 4 * D)
        self.zero_buf = torch.zeros(o_DX + K * N * D, device=self.dev)
        self._arena = (o_dZ, o_DQ, o_DX)
        self.flat_g = self.zero_buf[:n]
        self._dA = self.zero_buf[o_dA:o_dA + D * D].view(D, D)
        self._dZ = self.zero_buf[o_dZ:o_dZ + N * D].view(N, D)
        self._dq_stride = pad(N * 4 * D)
        self._DQKVA = [self.zero_buf[o_DQ + i * self._dq_stride:o_DQ + i * self._dq_stride + N * 4 * D].view(N, 4 * D)
                       for i in range(self.nDQ)]
        self._DX = self.zero_buf[o_DX:o_DX + K * N * D].view(K, N, D)      # dL/dx_k, accumulated by split-K
        self.adam_m = torch.zeros(n, device=self.dev)
        self.adam_v = torch.zeros(n, device=self.dev)
        self.p, self.g = [], []
        o = 0
        for p in self._slots:
            if p is None:                                  # slot absent in this head (NULL pointer at the C ABI)
                self.p.append(None)
                self.g.append(None)
                continue
            k = p.numel()
            if not reuse:
                self.flat_p[o:o + k].copy_(p.data.reshape(-1))
                p.data = self.flat_p[o:o + k].view(p.shape)
            p.grad = self.flat_g[o:o + k].view(p.shape)
            self.p.append(p.data)
            self.g.append(p.grad)
            o += k
        self.n_params = n
        # optimizer hyper-parameters live on the device so that a captured graph follows an lr scheduler
        self.hyper = torch.tensor([self.lr, self.betas[0], self.betas[1], self.eps, self.wd, 1.0], device=self.dev)
        self._opt = None
        self._opt_steps = 0

    def set_hyper(self, lr=None, betas=None, eps=None, weight_decay=None):
        if lr is not None:
            self.lr = float(lr)
        if betas is not None:
            self.betas = (float(betas[0]), float(betas[1]))
        if eps is not None:
            self.eps = float(eps)
        if weight_decay is not None:
            self.wd = float(weight_decay)
        self.hyper[:5].copy_(torch.tensor([self.lr, self.betas[0], self.betas[1], self.eps, self.wd]))

    def set_grad_scale(self, scale: float):
        """Device-side factor on dL/d(out) of the labelled heads (read by the captured loss kernel): 1/n when the losses
        of n steps are averaged before one optimizer step (train_sequence.py:57-60)."""
        self.hyper[5:6].copy_(torch.tensor([float(scale)]))

    # ------------------------------------------------------------------ torch.optim.Adam bridge
    def attach_optimizer(self, opt):
        """Make `opt` (torch.optim.Adam over model.parameters()) and the engine share ONE optimizer state: the
        engine's flat moment buffers become the optimizer's `exp_avg` / `exp_avg_sq` tensors (views), hyper-parameters
        are read from its param group, and the step count is kept in sync by pull_optimizer() / push_optimizer().
        So `optimizer.state_dict()` / `load_state_dict()` checkpoint and resume an engine-trained run exactly like
        the reference does (train_link.py:497-511, 587-603), and `optimizer.step()` on the module path continues the
        same moments."""
        if type(opt).__name__ != "Adam":
            raise NotImplementedError("StepEngine implements torch.optim.Adam")
        if len(opt.param_groups) != 1:
            raise NotImplementedError("one parameter group expected (torch.optim.Adam(model.parameters(), ...))")
        g = opt.param_groups[0]
        if g.get("amsgrad") or g.get("maximize") or g.get("capturable") or g.get("fused") or g.get("decoupled_weight_decay"):
            raise NotImplementedError("StepEngine's Adam kernel: amsgrad / maximize / capturable / fused must be off")
        self._opt = opt
        o = 0
        for p in self._plist:
            k = p.numel()
            st = opt.state.get(p) or {}
            if "exp_avg" in st:                              # adopt what the optimizer already holds (resumed run)
                self.adam_m[o:o + k].copy_(st["exp_avg"].reshape(-1))
                self.adam_v[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
            step = st.get("step", torch.tensor(0.0))
            opt.state[p] = {"step": step.detach().clone().float().cpu() if torch.is_tensor(step) else torch.tensor(float(step)),
                            "exp_avg": self.adam_m[o:o + k].view(p.shape),
                            "exp_avg_sq": self.adam_v[o:o + k].view(p.shape)}
            o += k
        self.pull_optimizer()

    def pull_optimizer(self):
        """optimizer -> engine: hyper-parameters (an lr scheduler may have changed them), step count, and moments that
        `load_state_dict` may have replaced."""
        opt = self._opt
        if opt is None:
            return
        g = opt.param_groups[0]
        self.set_hyper(g["lr"], g["betas"], g["eps"], g["weight_decay"])
        o, steps = 0, 0
        for p in self._plist:
            k = p.numel()
            st = opt.state.get(p)
            if st is None or "exp_avg" not in st:
                opt.state[p] = st = {"step": torch.tensor(0.0), "exp_avg": self.adam_m[o:o + k].view(p.shape),
                                     "exp_avg_sq": self.adam_v[o:o + k].view(p.shape)}
            elif st["exp_avg"].data_ptr() != self.adam_m.data_ptr() + 4 * o:     # load_state_dict made new tensors
                self.adam_m[o:o + k].copy_(st["exp_avg"].reshape(-1))
                self.adam_v[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
                st["exp_avg"] = self.adam_m[o:o + k].view(p.shape)
                st["exp_avg_sq"] = self.adam_v[o:o + k].view(p.shape)
            steps = max(steps, int(st["step"]))
            o += k
        self._opt_steps = steps
        self.w["counters"][2] = steps

    def push_optimizer(self):
        """engine -> optimizer: the step count after run(); the moments are shared storage already."""
        if self._opt is None:
            return
        for p in self._plist:
            st = self._opt.state[p]
            if torch.is_tensor(st["step"]):
                st["step"].fill_(float(self._opt_steps))
            else:
                st["step"] = torch.tensor(float(self._opt_steps))

    # ------------------------------------------------------------------ workspaces (upper bounds)
    def _alloc(self):
        dev, B, S, K, D = self.dev, self.B, self.S, self.K, self.D
        f32 = dict(device=dev, dtype=torch.float32)
        i64 = dict(device=dev, dtype=torch.long)
        n_seeds = self.n_seeds
        N, E, P = self.n_cap, self.e_cap, self.P
        w = {}
        w["counters"] = torch.zeros(8, **i64)
        w["frontier"] = torch.zeros(max(K - 1, 1) * N, **i64)
        w["frontier_cnt"] = torch.zeros(8, device=dev, dtype=torch.int32)
        # Front-end outputs (staged batch, seeds, scored pairs, sorted node set, CSR edge list, delta-t, global->local
        # map) exist twice: the stage/lookup/fill of step j+1 needs only the state written by step j's write-back,
        # which happens right after step j's forward, so it is issued beside step j's BACKWARD into the other set.
        self.fe = []
        for par in range(2):
            f = {}
            f["b_src"], f["b_dst"], f["b_t"] = (torch.zeros(B, **i64) for _ in range(3))
            f["seeds"] = torch.zeros(n_seeds, **i64)
            f["pair_src"], f["pair_dst"] = torch.zeros(P, **i64), torch.zeros(P, **i64)
            f["counts"] = torch.zeros(8, device=dev, dtype=torch.int32)
            f["n_id"] = torch.zeros(N, **i64)
            f["rowptr"] = torch.zeros(N + 1, device=dev, dtype=torch.int32)
            f["esrc"], f["edst"], f["eid"] = (torch.zeros(E, **i64) for _ in range(3))
            f["rel"] = torch.zeros(E, **f32)
            f["assoc"] = self.loader._assoc if par == 0 else torch.zeros_like(self.loader._assoc)
            f["Nd"], f["Ed"] = f["counts"][0:1], f["counts"][1:2]
            f["ctr"] = torch.zeros(8, **i64)              # [3] = step number of the batch staged in this set
            f["hub"] = torch.zeros(2 + min(N, 16384), dtype=torch.int32, device=dev)   # rows of degree > 8
            self.fe.append(f)
        w["counts"] = self.fe[0]["counts"]                 # (diagnostics)
        w["ins_ws"] = torch.zeros(4 * B, device=dev, dtype=torch.int32)
        w["z0"] = torch.zeros((N, D + self.Fn), **f32)
        w["X"] = torch.zeros((K + 1, N, D), **f32)
        w["EP"] = torch.zeros((E, D), **f32)
        self.hub_cap = min(N, 16384)                      # rows of degree > 8 go to the block-cooperative kernel
        w["A"] = torch.zeros((D, D), **f32)
        self.use_tc = _lib.use_tensor_cores(D)
        self._derive_choices()
        w["Wcat_hi"], w["Wcat_lo"] = torch.zeros((4 * D, D), **f32), torch.zeros((4 * D, D), **f32)
        self.use_tc_dx = (self.use_tc and D % 128 == 0          # dX contraction on tcgen05 needs 128-wide output tiles
                          and os.environ.get("CTAN_TC_DX", "1") == "1")
        w["WcatT_hi"], w["WcatT_lo"] = torch.zeros((D, 4 * D), **f32), torch.zeros((D, 4 * D), **f32)
        # no-grad steps: enc_x folded into the first iteration's projection (frozen weights inside a run(train=False)
        # call): one tcgen05 contraction of [memory row | node features | 0] against [Wenc; Wcat Wenc]
        self.fuse_enc = (self.use_tc and D % 128 == 0 and _lib.TC_MODE != 2 and not self.seq_mod
                         and os.environ.get("CTAN_INFER_FUSE_ENC", "1") == "1")
        # training steps (any num_iters: the first iteration): the same fold, the panel re-formed every step by ONE small
        # kernel on a side stream (ctan_fuse_enc_prep_split) -- one dependent stage fewer on the forward critical path
        self.fuse_train = (self.use_tc and D % 128 == 0 and _lib.TC_MODE != 2 and not self.seq_mod
                           and os.environ.get("CTAN_TRAIN_FUSE_ENC", "1") == "1")
        if self.fuse_enc or self.fuse_train:
            self.Kx = ((D + self.Fn + 31) // 32) * 32
            w["zext"] = torch.zeros((N + 128, self.Kx), **f32)
            w["Wce"], w["bce"] = torch.zeros((5 * D, self.Kx), **f32), torch.zeros(5 * D, **f32)
            w["Wce_hi"], w["Wce_lo"] = torch.zeros((5 * D, self.Kx), **f32), torch.zeros((5 * D, self.Kx), **f32)
            self._fuse_ver = None
        w["QKVA"] = torch.zeros((K, N, 4 * D), **f32)
        w["TH"] = torch.zeros((K, N, D), **f32)
        w["ALPHA"] = torch.zeros((K, E), **f32)
        w["Z"] = torch.zeros((N, D), **f32) if self.final_tanh else None
        w["H"] = torch.zeros((P, self.Hd), **f32)
        w["OUT"] = torch.zeros((P, self.O), **f32)
        w["dOUT"] = torch.zeros((P, self.O), **f32)
        w["loss_hist"] = torch.zeros(self.hist_cap, **f32)
        w["out_hist"] = torch.zeros((self.out_cap, P * self.O), **f32) if self.out_cap else None
        # backward
        w["dHid"] = torch.zeros((P, self.Hd), **f32)
        w["dZ"] = self._dZ
        w["gA"] = torch.zeros((N, D), **f32)
        w["DQKVA"] = self._DQKVA
        w["dEP"] = torch.zeros((E, D), **f32)
        w["DS"] = torch.zeros(E, **f32)
        w["dTenc"] = torch.zeros((E, max(self.T, 1)), **f32)
        w["dA"] = self._dA
        self.w = w
        self.prefetch = (not self.host_staged) and not self.seq_mod and os.environ.get("CTAN_TRAIN_PREFETCH", "1") == "1"
        # training steps too issue the neighbour insert and the next step's front-end (minus delta-t) at their first
        # instant, beside the forward: the state write-back + front-end branch had become as long as the backward chain
        self.early_train = os.environ.get("CTAN_TRAIN_EARLY", "1") == "1"

    # ------------------------------------------------------------------ state
    def reset(self, reset_optimizer: bool = False):
        """Fresh memory + empty graph + cursor at event 0 (train_link.py:226-227)."""
        self.model.reset_memory()
        self.loader.reset_state()
        # cursor / batch start / staged-step count / tickets restart; counters[2] is the optimizer's step count, which
        # torch.optim.Adam keeps across epochs (the reference resets only memory and graph, train_link.py:226-227)
        steps_done = self.w["counters"][2].clone()
        self.w["counters"].zero_()                           # (incl. [7], the event-id offset: the loader restarts at 0)
        if not reset_optimizer:
            self.w["counters"][2] = steps_done
        for f in self.fe:
            f["counts"].zero_()
        if self.host_staged:
            self._host_cursor = 0
            self._push_par = self._run_par = 0
            self._consumed = [None, None]
        if reset_optimizer:
            self.adam_m.zero_()
            self.adam_v.zero_()
            self._opt_steps = 0

    def push_batch(self, src, dst, t, neg, msg):
        """host_staged mode: copy one batch (pinned host tensors) to the device -- the ids/times into the
        staging buffers the step reads, the message rows and times into the event tables."""
        lo = self._host_cursor
        hi = lo + self.B
        par = self._push_par
        stg = self._stg[par]
        cs = self._copy_stream
        if self._consumed[par] is not None:
            cs.wait_event(self._consumed[par])               # the step that read this staging set has finished
        with torch.cuda.stream(cs):
            stg["src"].copy_(src, non_blocking=True)
            stg["dst"].copy_(dst, non_blocking=True)
            stg["t"].copy_(t, non_blocking=True)
            stg["neg"].copy_(neg.view(self.B, -1), non_blocking=True)
            self.msg[lo:hi].copy_(msg, non_blocking=True)
            self.t_tab[lo:hi].copy_(t, non_blocking=True)
            self._copied[par].record(cs)
        self._push_par = par ^ 1
        self._host_cursor = hi
        return self.B * (8 * (3 + self.n_neg) + 4 * self.Fe)

    def push_packed(self, ids, msg):
        """push_batch with the ids pre-packed by the data loader: `ids` pinned int64 [(3+n_neg)*B] laid out
        [src | dst | t | neg (event-major)], `msg` pinned [B,Fe].  Three copies instead of six (the host, not the GPU,
        bounds the end-to-end loop)."""
        lo = self._host_cursor
        hi = lo + self.B
        par = self._push_par
        stg = self._stg[par]
        cs = self._copy_stream
        if self._consumed[par] is not None:
            cs.wait_event(self._consumed[par])
        with torch.cuda.stream(cs):
            stg["ids"].copy_(ids, non_blocking=True)
            self.msg[lo:hi].copy_(msg, non_blocking=True)
            self.t_tab[lo:hi].copy_(stg["t"], non_blocking=True)
            self._copied[par].record(cs)
        self._push_par = par ^ 1
        self._host_cursor = hi
        return self.B * (8 * (3 + self.n_neg) + 4 * self.Fe)

    def set_cursor(self, event_index: int):
        self.w["counters"][0] = int(event_index)
        self.loader.cur_e_id = int(event_index)

    @property
    def cursor(self) -> int:
        return int(self.w["counters"][0].item())

    # ------------------------------------------------------------------ one step = one DAG of launches
    parallel = True                # False: issue everything on one stream (profiling / experiments)

    def _fork(self, i):
        if not self.parallel:
            return contextlib.nullcontext()
        self._side[i].wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(self._side[i])

    def _fork_from(self, i, ev):
        """Side stream i continues from the recorded point `ev` of the main stream (not from its tail): what the main
        stream issued after `ev` is a sibling of the side branch, created -- and so launched -- before it."""
        if not self.parallel:
            return contextlib.nullcontext()
        self._side[i].wait_event(ev)
        return torch.cuda.stream(self._side[i])

    def _join(self, i):
        if self.parallel:
            torch.cuda.current_stream().wait_stream(self._side[i])

    def _frontend(self, par: int, with_rel: bool = True):
        """stage -> K-hop lookup -> edge list (+ delta-t) of the batch at the cursor, into buffer set `par`.  Reads the
        neighbour tables and, with_rel, last_update: must follow the previous step's neighbour insert (and, with_rel,
        its memory write-back), nothing else."""
        w, m, ld, f = self.w, self.model, self.loader, self.fe[par]
        B, K, S = self.B, self.K, self.S
        N, E = self.n_cap, self.e_cap
        src, dst, t, neg = ((self._stg[par][k] for k in ("src", "dst", "t", "neg")) if self.host_staged
                            else (self.src, self.dst, self.t, self.neg))
        neg_p = None if neg is None else ptr(neg) - (0 if self.host_staged else 8 * self.n_neg * self.neg_base)
        call("ctan_stage_batch", ptr(src), ptr(dst), ptr(t), neg_p, self.n_neg, B,
             ptr(w["counters"]), 0 if self.host_staged else 1, ptr(f["b_src"]), ptr(f["b_dst"]), ptr(f["b_t"]),
             ptr(f["seeds"]), ptr(f["pair_src"]), ptr(f["pair_dst"]), ptr(w["loss_hist"]), self.hist_cap, ptr(f["ctr"]))
        call("ctan_nbr_lookup", ptr(ld.neighbors), ptr(ld._deg), self.num_nodes, S, ptr(f["seeds"]), self.n_seeds,
             None, K, ptr(ld._bitmap), ptr(ld._cbitmap), ptr(w["frontier"]), ptr(w["frontier_cnt"]), N,
             ptr(f["n_id"]), N, ptr(f["assoc"]), ptr(f["rowptr"]), E, ptr(f["counts"]))
        if with_rel and not self.seq_mod:
            call("ctan_nbr_fill_rel", ptr(ld.neighbors), ptr(ld.e_id), S, ptr(f["n_id"]), ptr(f["rowptr"]),
                 ptr(f["assoc"]), N, ptr(f["Nd"]), ptr(f["esrc"]), ptr(f["edst"]), ptr(f["eid"]), ptr(ld._bitmap),
                 ptr(ld._cbitmap), ptr(m.memory.last_update), ptr(self.t_tab), float(m.gnn.mean_delta_t),
                 float(m.gnn.std_delta_t), ptr(f["rel"]))
        else:
            call("ctan_nbr_fill", ptr(ld.neighbors), ptr(ld.e_id), S, ptr(f["n_id"]), ptr(f["rowptr"]), ptr(f["assoc"]),
                 N, ptr(f["Nd"]), ptr(f["esrc"]), ptr(f["edst"]), ptr(f["eid"]), ptr(ld._bitmap), ptr(ld._cbitmap))
        if self.seq_mod:                 # event id -> row of the current sequence's message / time table
            call("ctan_eid_mod", ptr(f["eid"]), E, ptr(f["Ed"]), self.seq_mod, ptr(f["ctr"]))
        if S > 8:                        # work list of the high-degree rows (edge_fwd runs them beside the plain rows)
            f["hub"][:2].zero_()
            call("ctan_hub_list", ptr(f["rowptr"]), N, ptr(f["Nd"]), ptr(f["hub"]), self.hub_cap)

    def _launch(self, train: bool):
        """One whole step, front-end inline (eager mode, warm-up, timeline scripts)."""
        try:
            self._prep(train)
            self._frontend(0)
            self._step(train, 0, False, prep_done=True)
        finally:
            _lib.lib().ctan_set_pdl(1, None)               # the launch-mode switch is thread-local host state: never leak it

    def _prep(self, train: bool, defer: bool = False):
        """Side branch with no dependence on the batch (weight-only operands, memsets) + the launch mode of the step.
        defer (folded training step whose front-end is already done): only the panel of the forward contraction is issued
        now; everything the forward does not need at once (_prep_rest) is issued by _step once the gather is under way,
        so that it does not compete with the two kernels the step's critical path starts with."""
        w, m = self.w, self.model
        D = self.D
        Wq, Wk, Wv, W = (ptr(self.p[i]) for i in (2, 4, 6, 9))
        if train and self.fuse_train:                        # side 3: [Wenc; Wcat Wenc] of THIS step's weights (joined just
            with self._fork(3):                              # before the forward contraction; the rest of the prep later)
                Wenc, benc, bq, bk, bv, b = (ptr(self.p[i]) for i in (0, 1, 3, 5, 7, 10))
                call("ctan_fuse_enc_prep_split", Wq, Wk, Wv, W, float(m.gnn.aconv.gamma), bq, bk, bv, b, Wenc, benc, D,
                     self.Fn, self.Kx, ptr(w["Wce_hi"]), ptr(w["Wce_lo"]), ptr(w["bce"]))
        if not defer:
            self._prep_rest(train)
        # Programmatic dependent launch (csrc/common.cuh) shortens the single-chain no-grad step (80 -> 75 us) but
        # LENGTHENS the training step (118 -> 122..127 us, more the further into the DAG it is enabled): dependents
        # scheduled early hold SM slots that the step's parallel branches need.  So: inference on, training off.
        _lib.lib().ctan_set_pdl(0 if (train and os.environ.get("CTAN_TRAIN_PDL", "0") != "1") else 1, None)

    def _prep_rest(self, train: bool):
        w, m = self.w, self.model
        D = self.D
        Wq, Wk, Wv, W = (ptr(self.p[i]) for i in (2, 4, 6, 9))
        with self._fork(2):
            if self.use_tc:      # A = W - W^T - gamma I, packed with q/k/v weights and split for 3xTF32
                call("ctan_wcat_prep", Wq, Wk, Wv, W, float(m.gnn.aconv.gamma), D, ptr(w["Wcat_hi"]),
                     ptr(w["Wcat_lo"]), ptr(w["A"]), ptr(w["WcatT_hi"]) if (train and self.use_tc_dx) else None,
                     ptr(w["WcatT_lo"]) if (train and self.use_tc_dx) else None)
            else:
                call("ctan_antisym_prep", W, float(m.gnn.aconv.gamma), D, ptr(w["A"]))
            w["H"].zero_()

    def _step(self, train: bool, par: int, prefetch: bool, prep_done: bool = False, early: bool = False):
        """Issue one step whose front-end is already in buffer set `par`.  Work that is off the critical path (edge
        projection, weight gradients, state write-back, and -- if `prefetch` -- the NEXT step's front-end into the
        other buffer set) goes to side streams with fork/join edges; under graph capture these become parallel
        branches of the captured DAG."""
        w, m, ld, f = self.w, self.model, self.loader, self.fe[par]
        B, D, K, S, Fe, Fn, T = self.B, self.D, self.K, self.S, self.Fe, self.Fn, self.T
        Hd, O, P = self.Hd, self.O, self.P
        N, E = self.n_cap, self.e_cap
        Nd, Ed = ptr(f["Nd"]), ptr(f["Ed"])
        assoc = f["assoc"]
        (Wenc, benc, Wq, bq, Wk, bk, Wv, bv, We, W, b, tw, tb, Wsrc, bsrc, Wdst, bdst, Wfin,
         bfin) = (ptr(t) for t in self.p)
        mem, lu = m.memory.memory, m.memory.last_update
        cur1 = ptr(w["counters"]) + 8                       # &counters[1]: current batch start
        xf = ptr(self.x) if Fn else None
        eps = float(m.gnn.aconv.epsilon)
        defer = (not prep_done) and train and self.fuse_train and os.environ.get("CTAN_DEFER_PREP", "1") == "1"
        if not prep_done:
            self._prep(train, defer)

        def backward_prep():                                 # the accumulation arena, cleared by the batch's TRUE node count,
            if defer:                                        # the saved copy of the gathered rows, and the deferred prep
                self._prep_rest(train)
            with self._fork(2):
                self._zero_arena(f)
            with self._fork(1):                              # side 1: saved copy of the gathered input rows
                call("ctan_gather_rows2", ptr(mem), D, xf, Fn, ptr(f["n_id"]), N, Nd, ptr(w["z0"]))
                self._z0_ready = torch.cuda.Event()
                self._z0_ready.record(torch.cuda.current_stream())
        if train and not defer:
            backward_prep()
        def insert():
            call("ctan_nbr_insert", ptr(ld.neighbors), ptr(ld.e_id), ptr(ld._deg), ptr(f["b_src"]), ptr(f["b_dst"]),
                 B, 0, cur1, S, ptr(w["ins_ws"]))
        if early:
            # no-grad pipelining: the neighbour insert needs only the staged batch, so it -- and then the NEXT step's
            # front-end (without delta-t, which reads this step's memory write-back) -- run beside the whole forward
            with self._fork(1):
                insert()
                if prefetch:
                    self._frontend(par ^ 1, with_rel=False)
        with self._fork(0):                                  # side 0: edge-attribute projection (|| enc_x)
            if early or self.seq_mod:                        # delta-t of THIS step (the front-end left it out)
                call("ctan_edge_rel", ptr(lu), ptr(f["n_id"]), ptr(f["esrc"]), ptr(self.t_tab), ptr(f["eid"]), E, Ed,
                     float(m.gnn.mean_delta_t), float(m.gnn.std_delta_t), ptr(f["rel"]))
            call(self._eproj_fn, ptr(self.msg), Fe, ptr(f["eid"]), ptr(f["rel"]), tw, tb, T, We, E, Ed, D,
                 ptr(w["EP"]))
        fused = self.fuse_train if train else self.fuse_enc
        if fused:                # X0 and QKVA of the first iteration in one contraction (panel: _prep / _prep_infer_weights)
            call("ctan_enc_prep_ext", ptr(mem), D, xf, Fn, ptr(f["n_id"]), N, Nd, self.Kx, ptr(w["zext"]))
            if defer:
                backward_prep()                              # (forked behind the gather: runs beside the contraction)
            if train:
                self._join(3)
            call("ctan_enc_qkva_tc", ptr(w["zext"]), N, Nd, self.Kx, ptr(w["Wce_hi"]), ptr(w["Wce_lo"]), D, ptr(w["bce"]),
                 ptr(w["X"][0]), ptr(w["QKVA"][0]), _lib.TC_MODE | self._row_hint)
        else:
            call("ctan_enc_fwd", ptr(mem), D, xf, Fn, ptr(f["n_id"]), None, N, Nd, Wenc, benc, ptr(w["X"][0]))
        late_join = fused and train      # side 2 (arena clear, Wcat panels, H) is first needed AFTER the first projection
        if not late_join:
            self._join(2)
        for k in range(K):
            xi = w["X"][k] if train else w["X"][k & 1]
            xo = w["X"][k + 1] if train else w["X"][(k + 1) & 1]
            qk = w["QKVA"][k] if train else w["QKVA"][0]
            if late_join and k == 1:
                self._join(2)    # (the Wcat panels)
            if fused and k == 0:
                pass             # (came with X0)
            elif self.use_tc:    # tcgen05 + TMA projection
                call("ctan_qkva_fwd_tc", ptr(xi), N, Nd, D, ptr(w["Wcat_hi"]), ptr(w["Wcat_lo"]), bq, bk, bv, b,
                     ptr(qk), _lib.TC_MODE | self._row_hint)
            else:
                call("ctan_qkva_fwd", ptr(xi), N, Nd, D, Wq, Wk, Wv, ptr(w["A"]), bq, bk, bv, b, ptr(qk))
            if k == 0:
                self._join(0)                                # edge projections (side 0) are first read by the edge kernel
            last = k == K - 1
            ef = (ptr(qk), ptr(w["EP"]), ptr(f["rowptr"]), ptr(f["esrc"]), N, Nd, D, ptr(xi), eps, ptr(xo),
                  ptr(w["TH"][k]) if train else None, ptr(w["ALPHA"][k]) if train else None,
                  ptr(w["Z"]) if (last and self.final_tanh) else None)
            if S > 8:
                with self._fork(0):
                    call("ctan_edge_fwd", *ef, ptr(f["hub"]), self.hub_cap, 2)         # hub rows (one CTA per row)
                call("ctan_edge_fwd", *ef, ptr(f["hub"]), self.hub_cap, 1)             # plain rows, concurrently
                self._join(0)
            else:
                call("ctan_edge_fwd", *ef, None, 0, 0)
        if late_join and K == 1:
            self._join(2)        # (H cleared for the readout; the arena and the Wcat^T panels for the backward)
        xk = w["X"][K] if train else w["X"][K & 1]
        z = w["Z"] if self.final_tanh else xk
        self._z = z

        def state_update():
            # ground-truth state update (train_link.py:276-277)
            call("ctan_mem_update", ptr(mem), ptr(lu), D, ptr(f["b_src"]), ptr(f["b_dst"]), ptr(f["b_t"]), B, None,
                 None, 0, ptr(z), ptr(assoc))
            if not early:
                insert()
        # no-grad step: the memory write-back and the readout both only READ z -- side by side (the next step's gather,
        # the first reader of the memory, is in the next graph)
        mem_side = (not train) and early and self.parallel and os.environ.get("CTAN_INFER_MEM_SIDE", "1") == "1"
        if mem_side:
            with self._fork(2):
                state_update()
        if self._link_fused:
            # hidden layer + head + BCE loss + dOUT + dHid + the dZ scatter in ONE launch (csrc/dense.cu::link_fused)
            call("ctan_readout_link_fused", ptr(z), D, ptr(f["pair_src"]), ptr(f["pair_dst"]), ptr(assoc), P, B, Hd,
                 Wsrc, bsrc, Wdst, bdst, Wfin, bfin, ptr(w["H"]), ptr(w["OUT"]), ptr(w["dOUT"]) if train else None,
                 ptr(w["dHid"]) if train else None, ptr(w["dZ"]) if train else None, ptr(w["loss_hist"]), self.hist_cap,
                 ptr(f["ctr"]))
        elif not self.labelled:
            call("ctan_readout_fwd", ptr(z), D, ptr(f["pair_src"]), ptr(f["pair_dst"]), ptr(assoc), P, None, Hd, O,
                 Wsrc, bsrc, Wdst, bdst, Wfin, bfin, ptr(w["H"]), None, 1)
            # head + BCE loss + dOUT + dHid in one launch (the loss slot was cleared by the stage kernel)
            call("ctan_head_loss", ptr(w["H"]), P, Hd, Wfin, bfin, B, ptr(w["OUT"]), ptr(w["dOUT"]) if train else None,
                 ptr(w["dHid"]) if train else None, ptr(w["loss_hist"]), self.hist_cap, ptr(f["ctr"]))
        else:
            # labelled heads (multi-class / regression / sequence): hidden + head, then loss(out, y) and dOUT
            call("ctan_readout_fwd", ptr(z), D, ptr(f["pair_src"]), ptr(f["pair_dst"]), ptr(assoc), P, None, Hd, O,
                 Wsrc, bsrc, Wdst, bdst, Wfin, bfin, ptr(w["H"]), ptr(w["OUT"]), 1)
            y_p = self.y.data_ptr() - self.y_base * (8 if self.loss_mode == 1 else 4 * O)
            call("ctan_loss_generic", ptr(w["OUT"]), P, O, self.loss_mode, y_p, float(self.grad_scale),
                 ptr(self.hyper) + 20, ptr(w["dOUT"]) if train else None, ptr(w["loss_hist"]), self.hist_cap, ptr(f["ctr"]))
        if w["out_hist"] is not None:
            call("ctan_record_rows", ptr(w["OUT"]), P * O, ptr(w["out_hist"]), self.out_cap, ptr(f["ctr"]))

        if not train:
            if mem_side:
                self._join(2)
            else:
                state_update()
            if early:
                self._join(1)
            _lib.lib().ctan_set_pdl(1, None)
            return
        with self._fork(1):                                  # side 1 (after the z0 gather): state write-back ...
            state_update()
            if prefetch and not early:                       # ... then the NEXT step's front-end, beside this backward
                self._frontend(par ^ 1)                      # (early: it already ran beside the forward, without delta-t)
        (gWenc, gbenc, gWq, gbq, gWk, gbk, gWv, gbv, gWe, gW, gb, gtw, gtb, gWsrc, gbsrc, gWdst, gbdst, gWfin,
         gbfin) = (ptr(t) for t in self.g)
        rb_args = (ptr(w["dOUT"]), ptr(w["H"]), ptr(z), D, ptr(f["pair_src"]), ptr(f["pair_dst"]), ptr(assoc), P,
                   None, Hd, O, Wsrc, Wdst, Wfin, ptr(w["dHid"]))
        ready = 0 if self.labelled else 1                    # (the fused binary head already produced dHid)
        if not self._link_fused:                             # (else dZ came with the forward launch)
            call("ctan_readout_bwd", *rb_args, ptr(w["dZ"]), None, None, None, None, None, None, ready)
        with self._fork(2):                                  # side 2: readout weight gradients
            call("ctan_readout_bwd", *rb_args, None, gWsrc, gbsrc, gWdst, gbdst, gWfin, gbfin, 1)
        g = w["dZ"]
        if self.final_tanh:
            call("ctan_tanh_bwd", ptr(w["dZ"]), ptr(z), N, Nd, D, ptr(w["gA"]))
            g = w["gA"]
        for k in range(K - 1, -1, -1):
            if k != K - 1 and self.nDQ == 1:
                self._join(0)                                # the previous iteration's dW contraction reads DQKVA
                self._zero_arena(f, only_dq=True)
            dq = w["DQKVA"][k if self.nDQ > 1 else 0]
            call("ctan_edge_bwd", ptr(w["QKVA"][k]), ptr(w["EP"]), ptr(f["rowptr"]), ptr(f["esrc"]), N, Nd, D, ptr(g),
                 ptr(w["TH"][k]), ptr(w["ALPHA"][k]), eps, ptr(dq), ptr(w["dEP"]), int(k != K - 1),
                 ptr(w["DS"]))
            qb = (ptr(dq), ptr(w["X"][k]), ptr(g), N, Nd, D, Wq, Wk, Wv, ptr(w["A"]))
            gn = self._DX[k]
            # dX is the critical path and its tcgen05 CTAs need a whole SM each: it is issued BEFORE the weight-gradient
            # branches that depend on the same edge_bwd, so that their many small CTAs do not take the SMs first
            dx_first = self.use_tc_dx and self.parallel and os.environ.get("CTAN_DX_FIRST", "1") == "1"
            def dx():
                if self.use_tc_dx:   # dX = G + DQKVA Wcat on tcgen05 (plain stores)
                    call("ctan_dx_tc", ptr(dq), ptr(g), N, Nd, D, ptr(w["WcatT_hi"]), ptr(w["WcatT_lo"]), ptr(gn),
                         _lib.TC_MODE | self._row_hint,
                         int(os.environ.get("CTAN_DX_SPLIT", "4")) if self.n_typ <= 4096 else 1)
                    # (kernel body at N=640: 18 / 12 / 9 us with 1 / 2 / 4 splits, scripts/micro_tc3.py)
                else:
                    call("ctan_qkva_bwd", *qb, ptr(gn), None, None, None, None, None, None, None, None, self.n_typ)
            if dx_first:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                dx()
                fork = lambda i: self._fork_from(i, ev)
            else:
                fork = self._fork
            with fork(0):                                    # side 0: weight gradients of this iteration
                if self._wgrad_direct:                       # (dA - dA^T accumulated straight into dW)
                    call("ctan_qkva_wgrad", ptr(dq), ptr(w["X"][k]), N, Nd, D, gWq, gWk, gWv, gW, gbq, gbk, gbv,
                         gb, self.n_typ)
                else:
                    call("ctan_qkva_bwd", *qb, None, gWq, gWk, gWv, ptr(w["dA"]), gbq, gbk, gbv, gb, self.n_typ)
                    if k == 0:
                        call("ctan_antisym_bwd", ptr(w["dA"]), D, gW)
            if k == 0:
                ep = (ptr(w["dEP"]), ptr(self.msg), Fe, ptr(f["eid"]), ptr(f["rel"]), tw, tb, T, We, E, Ed, D)
                with fork(3):                                # side 3: lin_edge weight gradient
                    call("ctan_eproj_bwd", *ep, gWe, None, self.e_typ)
                with fork(2):                                # side 2 (readout gradients long done): time-encoder gradients
                    if self._tenc_fused:
                        call("ctan_tenc_bwd_fused", ptr(w["dEP"]), We, Fe, T, D, ptr(f["rel"]), tw, tb, E, Ed, gtw, gtb)
                    else:
                        call("ctan_eproj_bwd", *ep, None, ptr(w["dTenc"]), self.e_typ)
                        call("ctan_tenc_bwd", ptr(w["dTenc"]), ptr(f["rel"]), tw, tb, E, Ed, T, gtw, gtb)
            if not dx_first:
                dx()
            g = gn
        if self.parallel:                                    # z0 is complete (side 1 itself -- state write-back, the next
            torch.cuda.current_stream().wait_event(self._z0_ready)    # step's front-end -- is only joined at the end)
        call("ctan_enc_bwd", ptr(g), ptr(w["z0"]), N, Nd, D, Fn, gWenc, gbenc, self.n_typ)
        self._join(0)
        self._join(2)
        self._join(3)
        if not self._accumulate:
            self._adam()
        self._join(1)
        _lib.lib().ctan_set_pdl(1, None)

    def _zero_arena(self, f, only_dq: bool = False):
        """Clear what the backward accumulates into: parameter gradients + dA (unless accumulating over steps) and the
        first N rows of dZ, DQKVA and dX_k -- N read on the device, not the workspace bound."""
        import ctypes
        N, D, K = self.n_cap, self.D, self.K
        o_dZ, o_DQ, o_DX = self._arena
        zb = self.zero_buf.data_ptr()
        if only_dq:
            segs = [(zb + 4 * o_DQ, 4 * D)]
            base, head = None, 0
        else:
            segs = ([(zb + 4 * o_dZ, D)] + [(zb + 4 * (o_DQ + i * self._dq_stride), 4 * D) for i in range(self.nDQ)]
                    + [(zb + 4 * (o_DX + k * N * D), D) for k in range(K)])
            base, head = (None, 0) if self._accumulate else (zb, o_dZ)
        for c0 in range(0, len(segs), 8):                    # (the kernel takes up to 8 segments per launch)
            part = segs[c0:c0 + 8]
            rows = (ctypes.c_int * 8)(*([r for _, r in part] + [0] * (8 - len(part))))
            ptrs = [p for p, _ in part] + [None] * (8 - len(part))
            call("ctan_zero_arena", base if c0 == 0 else None, head if c0 == 0 else 0, len(part), *ptrs,
                 ctypes.cast(rows, ctypes.c_void_p), ptr(f["Nd"]), N)

    def _prep_infer_weights(self):
        """[Wenc; Wcat Wenc] and its bias for the folded no-grad step; re-formed when the weights changed since the last
        no-grad run (torch tensor versions + _lib.WEIGHTS_EPOCH for the engine's own raw-pointer Adam)."""
        if not self.fuse_enc:
            return
        ver = (_lib.WEIGHTS_EPOCH,) + tuple((p.data_ptr(), p._version) for p in self._plist)
        if ver == self._fuse_ver:
            return
        w, m, D = self.w, self.model, self.D
        (Wenc, benc, Wq, bq, Wk, bk, Wv, bv, We, W, b) = (ptr(t) for t in self.p[:11])
        with torch.cuda.device(self.dev):
            call("ctan_antisym_prep", W, float(m.gnn.aconv.gamma), D, ptr(w["A"]))
            call("ctan_fuse_enc_prep", Wq, Wk, Wv, ptr(w["A"]), bq, bk, bv, b, Wenc, benc, D, self.Fn, self.Kx,
                 ptr(w["Wce"]), ptr(w["bce"]))
            call("ctan_pack_split", ptr(w["Wce"]), 5 * D, self.Kx, self.Kx, ptr(w["Wce_hi"]), ptr(w["Wce_lo"]))
        self._fuse_ver = ver

    def _adam(self):
        call("ctan_adam_step", ptr(self.flat_p), ptr(self.flat_g), ptr(self.adam_m), ptr(self.adam_v),
             self.n_params, self.lr, self.betas[0], self.betas[1], self.eps, self.wd, 0, ptr(self.w["counters"]),
             ptr(self.hyper))

    # ------------------------------------------------------------------ sequence mode: one step at a time
    def step(self, kind: str):
        """Sequence mode: one event.  "infer": no-grad step; "grad": forward + backward, gradients ACCUMULATED into the
        parameter gradients (scaled by set_grad_scale), no optimizer step."""
        key = ("seq", kind)
        g = self._graphs.get(key)
        if g is None:
            train = kind == "grad"
            self._accumulate = train
            try:
                if self.use_graph:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self._main):
                        self._prep(train)
                        self._frontend(0)
                        self._step(train, 0, False, prep_done=True)
                    g = g.replay
                else:
                    def g(train=train):
                        self._accumulate = train
                        try:
                            self._launch(train)
                        finally:
                            self._accumulate = False
            finally:
                self._accumulate = False
                _lib.lib().ctan_set_pdl(1, None)
            self._graphs[key] = g
        with torch.cuda.device(self.dev):
            g()
        self.loader.cur_e_id += self.B

    def adam(self):
        """Sequence mode: optimizer step on the accumulated gradients, then clear them (optimizer.zero_grad())."""
        g = self._graphs.get(("seq", "adam"))
        if g is None:
            def body():
                self._adam()
                self.zero_buf[:self._o_ws].zero_()
            if self.use_graph:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=self._main):
                    body()
                g = gr.replay
            else:
                g = body
            self._graphs[("seq", "adam")] = g
        _lib.WEIGHTS_EPOCH += 1
        with torch.cuda.device(self.dev):
            g()
        self._opt_steps += 1

    def zero_grads(self):
        """optimizer.zero_grad() for the accumulating steps: clear the parameter gradients (and dA)."""
        self.zero_buf[:self._o_ws].zero_()

    def set_eid_offset(self, offset: int):
        """Event id of stream row 0 (loader.cur_e_id at the start of a pass minus the cursor): the neighbour insert
        numbers the events of the pass from there."""
        self.w["counters"][7] = int(offset)

    # ------------------------------------------------------------------ public stepping API
    def _graph(self, train: bool, par: int = 0, prefetch: bool = False, inline: bool = True, early: bool = False):
        """Captured step.  inline: the front-end is part of the graph (one self-contained step; inference, host-staged
        mode).  Otherwise the front-end of buffer set `par` is expected to be done already and, if `prefetch`, the
        next step's front-end is a branch of this graph."""
        key = (bool(train), par, bool(prefetch), bool(inline), bool(early))
        if key not in self._graphs:
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g, stream=self._main):
                    if inline:
                        self._prep(train)
                        self._frontend(par)
                    self._step(train, par, prefetch, prep_done=inline, early=early)
            finally:
                _lib.lib().ctan_set_pdl(1, None)
            self._graphs[key] = g
        return self._graphs[key]

    def _graph_frontend(self, par: int = 0, with_rel: bool = True):
        key = ("fe", par, with_rel)
        if key not in self._graphs:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._main):
                self._frontend(par, with_rel)
            self._graphs[key] = g
        return self._graphs[key]

    def warmup(self):
        """Run one throw-away training + inference step eagerly (loads every kernel before graph
        capture) and restore the fresh state.  Call on a freshly reset engine only."""
        keep = self.flat_p.clone()
        base = max(self.neg_base, self.y_base)
        if int(self.w["counters"][0]) < base:                # the throw-away batch must lie inside the neg / label tables
            self.w["counters"][0] = base
        with torch.cuda.device(self.dev):
            self._launch(True)
            if base:
                self.w["counters"][0] = base
            self._prep_infer_weights()
            self._launch(False)
        torch.cuda.synchronize(self.dev)
        self.flat_p.copy_(keep)
        self.reset(reset_optimizer=True)
        self.w["loss_hist"].zero_()
        self._fuse_ver = None                                # (the panel above came from the throw-away step's weights)
        if not self.seq_mod:
            self.capture_all()

    def capture_all(self, train=(True, False)):
        """Capture every graph variant plan() can return, so that no capture ever lands in a caller's timed window."""
        if not self.use_graph:
            return
        with torch.cuda.device(self.dev):
            for tr in train:
                if self.host_staged:
                    for par in (0, 1):
                        self._graph(tr, par)
                elif not self.prefetch:
                    self._graph(tr)
                else:
                    early = (not tr) or self.early_train
                    self._graph_frontend(0, with_rel=not early)
                    for par in (0, 1):
                        for pf in (True, False):
                            self._graph(tr, par, pf, inline=False, early=early)
        torch.cuda.synchronize(self.dev)

    def plan(self, n_steps: int, train: bool = True):
        """The replay sequence of n_steps steps as a list of callables (one per step; the first also issues the first
        front-end).  Every step but the last of the sequence issues the next step's front-end beside its forward
        (CTAN_TRAIN_EARLY=0: training steps beside their backward): the state after the sequence is exactly that of
        n_steps self-contained steps."""
        if train:
            self._fuse_ver = None            # (training steps re-form the shared panel buffers and move the weights)
        else:
            self._prep_infer_weights()       # (the folded no-grad step reads a weight panel formed outside the graph)
        if self.host_staged:
            return [lambda: self._run_staged(train) for _ in range(n_steps)]
        if not self.use_graph:
            return [lambda: self._launch(train) for _ in range(n_steps)]
        if not self.prefetch:
            g = self._graph(train)
            return [g.replay for _ in range(n_steps)]
        seq, par = [], 0
        early = (not train) or self.early_train
        for j in range(n_steps):
            g = self._graph(train, par, j < n_steps - 1, inline=False, early=early)
            if j == 0:
                fe = self._graph_frontend(0, with_rel=not early)
                seq.append(lambda fe=fe, g=g: (fe.replay(), g.replay()))
            else:
                seq.append(g.replay)
            par ^= 1
        return seq

    def _run_staged(self, train: bool):
        """host_staged mode: one step on the batch of the oldest un-consumed push_batch()."""
        par = self._run_par
        cur = torch.cuda.current_stream()
        cur.wait_event(self._copied[par])
        if self.use_graph:
            self._graph(train, par).replay()
        else:
            self._prep(train)
            self._frontend(par)
            self._step(train, par, False, prep_done=True)
        ev = self._consumed_evs[par]                         # (re-recorded every second step; waits snapshot it)
        ev.record(cur)
        self._consumed[par] = ev
        self._run_par = par ^ 1

    def run(self, n_steps: int, train: bool = True):
        """Advance n_steps full batches from the cursor (no host synchronisation)."""
        if self.loader.cur_e_id + n_steps * self.B > self.n_events:
            raise ValueError(f"stream exhausted: cursor {self.loader.cur_e_id} + {n_steps}x{self.B} > {self.n_events}")
        if train:
            _lib.WEIGHTS_EPOCH += 1                          # (see _lib.WEIGHTS_EPOCH)
        with torch.cuda.device(self.dev):
            if self.host_staged:                             # (per-batch calls: keep the host path short)
                if train:
                    self._fuse_ver = None
                else:
                    self._prep_infer_weights()
                for _ in range(n_steps):
                    self._run_staged(train)
            else:
                for fn in self.plan(n_steps, train):
                    fn()
        self.loader.cur_e_id += n_steps * self.B
        if train:
            self._opt_steps += n_steps

    def outputs(self, n: int) -> torch.Tensor:
        """[n, P, O] scores of the last n steps (needs out_cap >= n; device tensor, reading synchronises).  Link head:
        rows [0,B) are the positives, [B, B(1+n_neg)) the negatives grouped by event."""
        if self.w["out_hist"] is None or n > self.out_cap:
            raise ValueError("construct the engine with out_cap >= n to keep per-step scores")
        steps = int(self.w["counters"][3].item())
        idx = torch.arange(steps - n, steps, device=self.dev) % self.out_cap
        return self.w["out_hist"][idx].view(n, self.P, self.O)

    def losses(self, n: int) -> torch.Tensor:
        """Last n recorded per-step losses (device tensor; reading it synchronises)."""
        steps = int(self.w["counters"][3].item())
        idx = torch.arange(steps - n, steps, device=self.dev) % self.hist_cap
        return self.w["loss_hist"][idx]

    def check(self):
        """Raise if any lookup overflowed its workspace during the run (device flag, one sync)."""
        err = max(int(f["counts"][3].item()) for f in self.fe)
        if err:
            raise _lib.CtanLibraryError(f"neighbor lookup overflowed the engine workspace (code {err})")
