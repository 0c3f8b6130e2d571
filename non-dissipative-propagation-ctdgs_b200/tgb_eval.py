"""TGB one-vs-many evaluation (tgb_test, split_mode 'val'/'test': train_link.py:111-159, 197-201) on our
kernels, restructured B200-first:

  * the reference runs one model forward PER QUERY (200 per batch); the memory and neighbour tables are
    frozen inside a batch (they are only updated at train_link.py:198-199), so node embeddings are
    query-independent: we run ONE forward per batch over the union of the seeds;
  * the queries of a batch are sharded over the ranks (rank r owns a contiguous slice); each rank scores
    its queries, then ONE all-gather per batch (NCCL over NVLink) exchanges, per query,
    [1/rank | scores | src embedding | dst embedding]; every rank then applies the identical memory
    write-back and neighbour insert, so the replicas stay bit-identical;
  * the MRR (TGB/tgb/linkproppred/evaluate.py:109-120) is computed on the device.

Embeddings are a deterministic function of a node's K-hop neighbourhood and are computed with the same
arithmetic order on every rank, so a sharded run reproduces the 1-GPU run exactly."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr
from .modules import CTAN_tgb
from .neighbor_loader import LastNeighborLoader


def shard_range(B: int, rank: int, world: int):
    """Contiguous, balanced slice [q0,q1) of a batch's B queries owned by `rank`."""
    base, rem = divmod(B, world)
    q0 = rank * base + min(rank, rem)
    return q0, q0 + base + (1 if rank < rem else 0)


def exchange_rows(send: torch.Tensor, recv: torch.Tensor, rows: torch.Tensor, B: int, world: int, group=None):
    """The single collective of a batch: all-gather every rank's packed query rows (`send`
    [Qmax,row], padded) into `recv` [world*Qmax,row], then compact the per-rank slices (shards may differ
    by one query) into batch order `rows` [B,row].  Works on any backend (NCCL on GPU, gloo in CPU tests)."""
    import torch.distributed as dist
    qmax = send.size(0)
    if B % world == 0:                 # equal shards: the gathered buffer already is the batch order
        dist.all_gather_into_tensor(rows, send, group=group)
        return rows
    dist.all_gather_into_tensor(recv, send, group=group)
    for rk in range(world):
        a0, a1 = shard_range(B, rk, world)
        if a1 > a0:
            rows[a0:a1].copy_(recv[rk * qmax: rk * qmax + (a1 - a0)])
    return rows


class TGBEvalEngine:
    def __init__(self, model: CTAN_tgb, loader: LastNeighborLoader, stream: dict, batch_size: int, n_neg: int,
                 rank: int = 0, world: int = 1, group=None, use_graph: bool = True, neg_base: int = 0,
                 exchange: str = "auto"):
        """neg_base: event index of row 0 of stream["neg"] when the negatives cover only a suffix of the event stream
        (a validation / test split: train_link.py:117 queries them per split)."""
        dev = model.memory.memory.device
        if dev.type != "cuda":
            raise _lib.CtanLibraryError("TGBEvalEngine needs the model on a CUDA device")
        _lib.lib()
        self.model, self.loader, self.dev = model, loader, dev
        self.rank, self.world, self.group, self.use_graph = rank, world, group, use_graph
        self.B, self.n_neg = int(batch_size), int(n_neg)
        self.q0, self.q1 = shard_range(self.B, rank, world)
        self.Q = self.q1 - self.q0
        self.Qmax = shard_range(self.B, 0, world)[1]                 # largest shard (all-gather row count)
        self.K, self.D, self.S = model.num_gnn_layers, model.memory.memory_dim, loader.size
        self.num_nodes = model.num_nodes
        self.final_tanh = model.final_act == "tanh"
        st = stream
        for k in ("src", "dst", "t", "msg", "neg"):
            if not st[k].is_cuda:
                raise _lib.CtanLibraryError(f"stream['{k}'] must be device resident")
        self.src, self.dst, self.t = st["src"].contiguous(), st["dst"].contiguous(), st["t"].contiguous()
        self.msg = st["msg"].float().contiguous()
        self.n_events = self.src.numel()
        self.neg = st["neg"].contiguous().view(-1, self.n_neg)
        self.neg_base = int(neg_base)
        if self.neg.size(0) + self.neg_base < self.n_events and self.neg_base == 0:
            raise ValueError("stream['neg'] must cover every event (or pass neg_base)")
        self.x = st.get("x")
        if self.x is not None:
            self.x = (self.x.squeeze(0) if self.x.dim() == 3 else self.x).to(dev).float().contiguous()
        self.Fe, self.Fn = self.msg.size(1), (self.x.size(1) if self.x is not None else 0)
        self.T = model.gnn.time_enc.lin.bias.numel()
        r = model.readout
        self.Hd, self.O = r.lin_src.weight.size(0), r.lin_final.weight.size(0)
        assert self.O == 1
        self.L = 1 + self.n_neg
        self.P = max(self.Q, 1) * self.L
        self.n_seeds = max(self.Q, 1) * (2 + self.n_neg)
        self.n_cap = min(self.num_nodes, self.n_seeds * (self.S + 1) ** self.K)
        c_cap = min(self.num_nodes, self.n_seeds * (self.S + 1) ** (self.K - 1))
        self.e_cap = max(c_cap * self.S, 1)
        self.row = 1 + self.L + 2 * self.D                          # packed per-query row
        self._alloc()
        self._graph = None
        import os
        # The next batch's front-end (stage, K-hop lookup, fill) only needs this batch's neighbour insert, so it is issued
        # ahead, into the other buffer set.  One rank ("1"): beside the scoring (125 -> 115 us per batch).  Sharded ("2"):
        # forked AFTER the scoring chain, beside the exchange -- push -> wait -> update is mostly waiting on NVLink, while
        # beside the scoring its many small CTAs keep the tcgen05 CTAs (a whole SM each) off the SMs: per batch at
        # 2 ranks 111.8 inline / 104.6 beside the scoring / 101.3 us beside the exchange; at 8 ranks 91.8 / 92.8 / 81.8 us.
        pf = os.environ.get("CTAN_EVAL_PREFETCH", "1" if world == 1 else "2")
        self.prefetch = pf in ("1", "2")
        # "2": the next front-end is forked AFTER the scoring chain, beside the exchange (push -> wait -> update is mostly
        # waiting on NVLink), instead of beside the scoring, whose tcgen05 CTAs it would keep off the SMs
        self.prefetch_late = pf == "2"
        self._side = torch.cuda.Stream(device=dev)
        self._side2 = torch.cuda.Stream(device=dev)
        # the per-batch exchange: "peer" = our kernels store the packed rows into every peer's symmetric buffer over
        # NVLink from INSIDE the captured graph (csrc/sym.cu); "nccl" = torch's all_gather_into_tensor between two graphs
        import os as _os
        if exchange == "auto":
            exchange = _os.environ.get("CTAN_EVAL_EXCHANGE", "peer")
        self.exchange = exchange if world > 1 else "none"
        if self.exchange == "nccl" and "CTAN_EVAL_PREFETCH" not in _os.environ:
            # (two graphs around a host-issued collective: no place beside the exchange; beside the scoring only for big shards)
            self.prefetch, self.prefetch_late = self.n_seeds >= 1000, False
        self.sym = None
        if self.exchange == "peer":
            from .sym import SymmetricExchange
            self.sym = SymmetricExchange(self.B, self.row, rank, world, dev, group)
        self.exchange_kind = {"none": "none (1 rank)", "nccl": "NCCL all_gather_into_tensor between two CUDA graphs",
                              "peer": "in-graph peer-memory stores over NVLink (ctan_xchg_push / ctan_xchg_wait)"}[self.exchange]

    def _alloc(self):
        dev, B, K, D = self.dev, self.B, self.K, self.D
        N, E, P = self.n_cap, self.e_cap, self.P
        f32 = dict(device=dev, dtype=torch.float32)
        i64 = dict(device=dev, dtype=torch.long)
        w = {}
        w["counters"] = torch.zeros(8, **i64)
        w["frontier"] = torch.zeros(max(K - 1, 1) * N, **i64)
        w["frontier_cnt"] = torch.zeros(8, device=dev, dtype=torch.int32)
        # Front-end outputs (staged batch, seeds, scored pairs, sorted node set, CSR edge list, global->local map)
        # are double-buffered: the stage/lookup/fill of batch j+1 depends only on the neighbour tables (final once
        # batch j's insert is done), so it runs as a side branch WHILE batch j is scored -- off the critical path.
        self.fe = []
        for par in range(2):
            f = {}
            f["b_src"], f["b_dst"], f["b_t"] = (torch.zeros(B, **i64) for _ in range(3))
            f["seeds"] = torch.zeros(self.n_seeds, **i64)
            f["pair_src"], f["pair_dst"] = torch.zeros(P, **i64), torch.zeros(P, **i64)
            f["counts"] = torch.zeros(8, device=dev, dtype=torch.int32)
            f["n_id"] = torch.zeros(N, **i64)
            f["rowptr"] = torch.zeros(N + 1, device=dev, dtype=torch.int32)
            f["esrc"], f["edst"], f["eid"] = (torch.zeros(E, **i64) for _ in range(3))
            f["assoc"] = self.loader._assoc if par == 0 else torch.zeros_like(self.loader._assoc)
            f["hub"] = torch.zeros(2 + min(N, 16384), dtype=torch.int32, device=dev)   # rows of degree > 8 (ctan_hub_list)
            f["Nd"], f["Ed"] = f["counts"][0:1], f["counts"][1:2]
            self.fe.append(f)
        w["counts"] = self.fe[0]["counts"]                 # (diagnostics: sizes of an even batch's subgraph)
        w["ins_ws"] = torch.zeros(4 * B, device=dev, dtype=torch.int32)
        w["rel"] = torch.zeros(E, **f32)
        w["X"] = torch.zeros((2, N, D), **f32)
        w["EP"] = torch.zeros((E, D), **f32)
        self.hub_cap = min(N, 16384)                      # rows of degree > 8 go to the block-cooperative kernel
        w["A"] = torch.zeros((D, D), **f32)
        self.use_tc = _lib.use_tensor_cores(D)
        # precision of the tensor-core contractions: 0 = 3xTF32 (fp32 parity), 1 = TF32, 2 = BF16 (needs D % 64 == 0)
        self.tc_mode = _lib.TC_MODE if (_lib.TC_MODE != 2 or D % 64 == 0) else 0
        # expected node / edge row counts, packed into bits 8+ of the tcgen05 entry points' `mode` argument (their CTAs
        # need a whole SM each: the persistent grid is sized by the typical subgraph, not by the workspace bound)
        import os as _os
        hint = _os.environ.get("CTAN_TC_ROW_HINT", "1") == "1"
        self._n_hint = (min(N, 2 * self.n_seeds) << 8) if hint else 0
        self._e_hint = (min(E, 4 * self.n_seeds) << 8) if hint else 0
        bf16 = dict(device=dev, dtype=torch.bfloat16)
        # tensor-core enc_x (dense gathered rows + packed weights) and node-projection readout
        self.use_tc_enc = self.use_tc and D % 128 == 0
        self.use_tc_ro = self.use_tc and (2 * self.Hd) % 128 == 0 and self.Hd % 16 == 0
        self.use_tc_ep = self.use_tc and D % 128 == 0
        if self.use_tc_ep:               # lin_edge on the tensor cores: dense [E, Kpad] edge_attr operand
            kq = 64 if self.tc_mode == 2 else 32
            self.Kep = ((self.Fe + self.T + kq - 1) // kq) * kq
            w["Wep_bf"] = torch.zeros((D, self.Kep), **bf16)
            w["Aep"] = torch.zeros((self.e_cap + 128, self.Kep), **f32)
            w["Wep_hi"], w["Wep_lo"] = torch.zeros((D, self.Kep), **f32), torch.zeros((D, self.Kep), **f32)
        # num_iters = 1 with frozen weights: enc_x and the projection are one contraction against [Wenc; Wcat Wenc]
        # (csrc/tc_gemm.cu::ctan_enc_qkva_tc) -- one tcgen05 launch and one dependent stage fewer per batch
        self.fuse_enc = (self.use_tc_enc and K == 1 and _os.environ.get("CTAN_EVAL_FUSE_ENC", "1") == "1")
        if self.fuse_enc:
            kq = 64 if self.tc_mode == 2 else 32
            self.Kx = ((D + self.Fn + kq - 1) // kq) * kq
            w["zext"] = torch.zeros((N + 128, self.Kx), **f32)
            w["Wce"], w["bce"] = torch.zeros((5 * D, self.Kx), **f32), torch.zeros(5 * D, **f32)
            w["Wce_hi"], w["Wce_lo"] = torch.zeros((5 * D, self.Kx), **f32), torch.zeros((5 * D, self.Kx), **f32)
            w["Wce_bf"] = torch.zeros((5 * D, self.Kx), **bf16)
        if self.use_tc_enc:
            w["zmain"], w["ztail"] = torch.zeros((N, D), **f32), torch.zeros((N, D), **f32)
            w["Wenc_hi"], w["Wenc_lo"] = torch.zeros((D, D), **f32), torch.zeros((D, D), **f32)
            w["Wenc_bf"] = torch.zeros((D, D), **bf16)
        if self.use_tc_ro:
            w["Wro_hi"], w["Wro_lo"] = torch.zeros((2 * self.Hd, D), **f32), torch.zeros((2 * self.Hd, D), **f32)
            w["Y"] = torch.zeros((N, 2 * self.Hd), **f32)
            w["Wro_bf"] = torch.zeros((2 * self.Hd, D), **bf16)
        w["Wcat_hi"], w["Wcat_lo"] = torch.zeros((4 * D, D), **f32), torch.zeros((4 * D, D), **f32)
        w["Wcat_bf"] = torch.zeros((4 * D, D), **bf16)
        w["QKVA"] = torch.zeros((N, 4 * D), **f32)
        w["Z"] = torch.zeros((N, D), **f32) if self.final_tanh else None
        w["H"] = torch.zeros((P, self.Hd), **f32)
        w["OUT"] = torch.zeros((P, self.O), **f32)
        w["send"] = torch.zeros((self.Qmax, self.row), **f32)
        w["recv"] = torch.zeros((self.world * self.Qmax, self.row), **f32) if self.world > 1 else w["send"]
        w["rows"] = torch.zeros((B, self.row), **f32) if self.world > 1 else w["send"]
        w["acc"] = torch.zeros(2, device=dev, dtype=torch.float64)   # [sum of 1/rank, #queries]
        self.w = w

    # ------------------------------------------------------------------ state
    def set_cursor(self, event_index: int):
        self.w["counters"][0] = int(event_index)
        self.loader.cur_e_id = int(event_index)

    def reset_metric(self):
        self.w["acc"].zero_()

    def mrr(self) -> float:
        a = self.w["acc"].cpu()
        return float(a[0] / a[1]) if float(a[1]) > 0 else float("nan")

    # ------------------------------------------------------------------ one batch
    def _launch(self, dry: bool = False):
        self._frontend(0)
        self._batch(0, False, dry)
        xpar = self._exchange()
        self._update(0, dry, xpar)

    def _frontend(self, par: int):
        """stage -> K-hop lookup -> edge list of the batch at the cursor, into buffer set `par`.  Reads the neighbour
        tables only (not the memory): may run before the previous batch's memory write-back."""
        w, ld, f = self.w, self.loader, self.fe[par]
        B, K, S, Q = self.B, self.K, self.S, self.Q
        N, E = self.n_cap, self.e_cap
        call("ctan_stage_eval", ptr(self.src), ptr(self.dst), ptr(self.t),
             ptr(self.neg) - 8 * self.n_neg * self.neg_base, self.n_neg, B, self.q0,
             self.q1, ptr(w["counters"]), ptr(f["b_src"]), ptr(f["b_dst"]), ptr(f["b_t"]), ptr(f["seeds"]),
             ptr(f["pair_src"]), ptr(f["pair_dst"]))
        if Q > 0:
            call("ctan_nbr_lookup", ptr(ld.neighbors), ptr(ld._deg), self.num_nodes, S, ptr(f["seeds"]),
                 Q * (2 + self.n_neg), None, K, ptr(ld._bitmap), ptr(ld._cbitmap), ptr(w["frontier"]),
                 ptr(w["frontier_cnt"]), N, ptr(f["n_id"]), N, ptr(f["assoc"]), ptr(f["rowptr"]), E, ptr(f["counts"]))
            call("ctan_nbr_fill", ptr(ld.neighbors), ptr(ld.e_id), S, ptr(f["n_id"]), ptr(f["rowptr"]), ptr(f["assoc"]),
                 N, ptr(f["Nd"]), ptr(f["esrc"]), ptr(f["edst"]), ptr(f["eid"]), ptr(ld._bitmap), ptr(ld._cbitmap))
            if S > 8:                    # work list of the high-degree rows: plain and hub rows then run concurrently
                f["hub"][:2].zero_()
                call("ctan_hub_list", ptr(f["rowptr"]), N, ptr(f["Nd"]), ptr(f["hub"]), self.hub_cap)

    def _insert(self, par: int, dry: bool):
        """Neighbour-table insert of the batch in buffer set `par` (needs only the staged (src, dst); must follow that
        batch's lookup/fill, the last readers of the tables)."""
        w, ld, f = self.w, self.loader, self.fe[par]
        if dry:
            sm, sl, sn, se, sd = self._scratch
            call("ctan_nbr_insert", ptr(sn), ptr(se), ptr(sd), ptr(sl), ptr(sl), 1, 0, None, self.S, ptr(w["ins_ws"]))
        else:
            call("ctan_nbr_insert", ptr(ld.neighbors), ptr(ld.e_id), ptr(ld._deg), ptr(f["b_src"]), ptr(f["b_dst"]),
                 self.B, 0, ptr(w["counters"]) + 8, self.S, ptr(w["ins_ws"]))

    def _batch(self, par: int, prefetch: bool, dry: bool = False):
        """Score the batch whose front-end is in buffer set `par`; beside it (side branch): its neighbour insert and,
        if `prefetch`, the front-end of the NEXT batch into the other buffer set."""
        cur = torch.cuda.current_stream()
        late = prefetch and self.prefetch_late and self.exchange != "nccl"
        self._side2.wait_stream(cur)
        with torch.cuda.stream(self._side2):
            self._insert(par, dry)            # reads counters[1] of this batch: before the next stage overwrites it
            if prefetch and not late:
                self._frontend(par ^ 1)
        self._forward(par)
        if late:                              # same stream as the insert: ordered behind it; joined by _join_late()
            self._side2.wait_stream(cur)
            with torch.cuda.stream(self._side2):
                self._frontend(par ^ 1)
            self._late_pending = True
        else:
            cur.wait_stream(self._side2)

    def _join_late(self):
        """End of a batch whose next front-end was forked beside the exchange (prefetch_late)."""
        if getattr(self, "_late_pending", False):
            torch.cuda.current_stream().wait_stream(self._side2)
            self._late_pending = False

    def _panel(self, name: str):
        """(B_hi, B_lo) pointers of a weight panel for the current precision mode (BF16: one panel, no lo part)."""
        if self.tc_mode == 2:
            return self.w[name + "_bf"].data_ptr(), None
        return ptr(self.w[name + "_hi"]), ptr(self.w[name + "_lo"])

    def _forward(self, par: int):
        """delta-t -> forward -> scores -> packed per-query rows of this rank's shard."""
        w, m, f = self.w, self.model, self.fe[par]
        B, D, K, S, Fe, Fn, T, Hd, O, P, Q = self.B, self.D, self.K, self.S, self.Fe, self.Fn, self.T, self.Hd, self.O, self.P, self.Q
        N, E = self.n_cap, self.e_cap
        Nd, Ed = ptr(f["Nd"]), ptr(f["Ed"])
        g, a, r = m.gnn, m.gnn.aconv, m.readout
        ph = a.phi
        mem, lu = m.memory.memory, m.memory.last_update
        xf = ptr(self.x) if Fn else None
        eps = float(a.epsilon)
        if Q > 0:
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                # delta-t reads last_update, i.e. the memory write-back of the previous batch: it belongs here, not
                # to the (prefetched) front-end
                call("ctan_edge_rel", ptr(lu), ptr(f["n_id"]), ptr(f["esrc"]), ptr(self.t), ptr(f["eid"]), E, Ed,
                     float(g.mean_delta_t), float(g.std_delta_t), ptr(w["rel"]))
                if self.use_tc_ep:
                    call("ctan_eproj_prep", ptr(self.msg), Fe, ptr(f["eid"]), ptr(w["rel"]), ptr(g.time_enc.lin.weight),
                         ptr(g.time_enc.lin.bias), T, E, Ed, self.Kep, ptr(w["Aep"]))
                    call("ctan_gemm_tc", ptr(w["Aep"]), E, Ed, self.Kep, *self._panel("Wep"), D, None,
                         None, None, None, D, None, 0, ptr(w["EP"]), self.tc_mode | self._e_hint)
                else:
                    call("ctan_eproj_fwd", ptr(self.msg), Fe, ptr(f["eid"]), ptr(w["rel"]), ptr(g.time_enc.lin.weight),
                         ptr(g.time_enc.lin.bias), T, ptr(ph.lin_edge.weight), E, Ed, D, ptr(w["EP"]))
            if self.fuse_enc:            # gather -> [mem | x | 0] rows; enc_x AND q/k/v/x.A^T in one tcgen05 contraction
                call("ctan_enc_prep_ext", ptr(mem), D, xf, Fn, ptr(f["n_id"]), N, Nd, self.Kx, ptr(w["zext"]))
                call("ctan_enc_qkva_tc", ptr(w["zext"]), N, Nd, self.Kx, *self._panel("Wce"), D, ptr(w["bce"]),
                     ptr(w["X"][0]), ptr(w["QKVA"]), self.tc_mode | self._n_hint)
            elif self.use_tc_enc:        # gather -> dense rows; enc_x on the tensor cores; node-feature columns as residual
                call("ctan_enc_prep", ptr(mem), D, xf, Fn, ptr(f["n_id"]), N, Nd, ptr(g.enc_x.weight), ptr(w["zmain"]),
                     ptr(w["ztail"]))
                call("ctan_gemm_tc", ptr(w["zmain"]), N, Nd, D, *self._panel("Wenc"), D,
                     ptr(g.enc_x.bias), None, None, None, D, ptr(w["ztail"]), D, ptr(w["X"][0]), self.tc_mode | self._n_hint)
            else:
                call("ctan_enc_fwd", ptr(mem), D, xf, Fn, ptr(f["n_id"]), None, N, Nd, ptr(g.enc_x.weight),
                     ptr(g.enc_x.bias), ptr(w["X"][0]))
            torch.cuda.current_stream().wait_stream(self._side)
            for k in range(K):
                xi, xo = w["X"][k & 1], w["X"][(k + 1) & 1]
                if self.fuse_enc and k == 0:
                    pass                                                 # QKVA of the only iteration came with X0
                elif self.use_tc:
                    call("ctan_qkva_fwd_tc", ptr(xi), N, Nd, D, *self._panel("Wcat"),
                         ptr(ph.lin_query.bias), ptr(ph.lin_key.bias), ptr(ph.lin_value.bias), ptr(a.bias),
                         ptr(w["QKVA"]), self.tc_mode | self._n_hint)
                else:
                    call("ctan_qkva_fwd", ptr(xi), N, Nd, D, ptr(ph.lin_query.weight), ptr(ph.lin_key.weight),
                         ptr(ph.lin_value.weight), ptr(w["A"]), ptr(ph.lin_query.bias), ptr(ph.lin_key.bias),
                         ptr(ph.lin_value.bias), ptr(a.bias), ptr(w["QKVA"]))
                last = k == K - 1
                ef = (ptr(w["QKVA"]), ptr(w["EP"]), ptr(f["rowptr"]), ptr(f["esrc"]), N, Nd, D, ptr(xi), eps, ptr(xo), None,
                      None, ptr(w["Z"]) if (last and self.final_tanh) else None)
                if self.S > 8:
                    self._side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(self._side):
                        call("ctan_edge_fwd", *ef, ptr(f["hub"]), self.hub_cap, 2)     # hub rows (one CTA per row)
                    call("ctan_edge_fwd", *ef, ptr(f["hub"]), self.hub_cap, 1)         # plain rows, concurrently
                    torch.cuda.current_stream().wait_stream(self._side)
                else:
                    call("ctan_edge_fwd", *ef, None, 0, 0)
            z = w["Z"] if self.final_tanh else w["X"][K & 1]
            if self.use_tc_ro:
                # project every NODE once (tensor cores), then a scored pair only gathers two rows
                call("ctan_gemm_tc", ptr(z), N, Nd, D, *self._panel("Wro"), 2 * Hd, ptr(r.lin_src.bias),
                     ptr(r.lin_dst.bias), None, None, Hd, None, 0, ptr(w["Y"]), self.tc_mode | self._n_hint)
                call("ctan_head_pack", ptr(w["Y"]), Hd, ptr(z), D, ptr(f["assoc"]), ptr(f["pair_src"]), ptr(f["pair_dst"]),
                     Q, self.n_neg, ptr(r.lin_final.weight), ptr(r.lin_final.bias), ptr(w["send"]))
            else:
                call("ctan_readout_fwd", ptr(z), D, ptr(f["pair_src"]), ptr(f["pair_dst"]), ptr(f["assoc"]), Q * self.L,
                     None, Hd, O, ptr(r.lin_src.weight), ptr(r.lin_src.bias), ptr(r.lin_dst.weight), ptr(r.lin_dst.bias),
                     ptr(r.lin_final.weight), ptr(r.lin_final.bias), ptr(w["H"]), ptr(w["OUT"]), 0)   # no split-K: scores must not depend on atomic order
                call("ctan_pack_eval", ptr(w["OUT"]), ptr(z), ptr(f["assoc"]), ptr(f["pair_src"]), ptr(f["pair_dst"]), Q,
                     self.n_neg, D, ptr(w["send"]))

    def _prep_weights(self):
        """Weight-only operands (A = W - W^T - gamma I, the packed / hi-lo split weight panels of the tensor-core
        contractions).  Evaluation runs with frozen weights (tgb_test is @torch.no_grad, train_link.py:111),
        so they are formed when the weights have changed since the last run() call (torch tensor versions +
        _lib.WEIGHTS_EPOCH for raw-pointer writers; `refresh_weights()` forces it), never once per batch."""
        w, m = self.w, self.model
        D, Fn, Hd = self.D, self.Fn, self.Hd
        g, a, r = m.gnn, m.gnn.aconv, m.readout
        ph = a.phi
        with torch.cuda.device(self.dev):
            if self.use_tc:
                call("ctan_wcat_prep", ptr(ph.lin_query.weight), ptr(ph.lin_key.weight), ptr(ph.lin_value.weight),
                     ptr(a.W), float(a.gamma), D, ptr(w["Wcat_hi"]), ptr(w["Wcat_lo"]), ptr(w["A"]), None, None)
            else:
                call("ctan_antisym_prep", ptr(a.W), float(a.gamma), D, ptr(w["A"]))
            if self.use_tc_ro:           # [lin_src; lin_dst] packed + split for the node-projection readout
                call("ctan_pack_split", ptr(r.lin_src.weight), Hd, D, D, ptr(w["Wro_hi"]), ptr(w["Wro_lo"]))
                call("ctan_pack_split", ptr(r.lin_dst.weight), Hd, D, D, ptr(w["Wro_hi"]) + 4 * Hd * D,
                     ptr(w["Wro_lo"]) + 4 * Hd * D)
            if self.use_tc_enc:
                call("ctan_pack_split", ptr(g.enc_x.weight), D, D, D + Fn, ptr(w["Wenc_hi"]), ptr(w["Wenc_lo"]))
            if self.use_tc_ep:
                call("ctan_pack_split_pad", ptr(ph.lin_edge.weight), D, self.Fe + self.T, self.Fe + self.T, self.Kep,
                     ptr(w["Wep_hi"]), ptr(w["Wep_lo"]))
            if self.fuse_enc:            # [Wenc; Wcat Wenc] and its bias, then the hi/lo (or BF16) panel of it
                call("ctan_fuse_enc_prep", ptr(ph.lin_query.weight), ptr(ph.lin_key.weight), ptr(ph.lin_value.weight),
                     ptr(w["A"]), ptr(ph.lin_query.bias), ptr(ph.lin_key.bias), ptr(ph.lin_value.bias), ptr(a.bias),
                     ptr(g.enc_x.weight), ptr(g.enc_x.bias), D, Fn, self.Kx, ptr(w["Wce"]), ptr(w["bce"]))
                call("ctan_pack_split", ptr(w["Wce"]), 5 * D, self.Kx, self.Kx, ptr(w["Wce_hi"]), ptr(w["Wce_lo"]))
                if self.tc_mode == 2:
                    call("ctan_pack_bf16", ptr(w["Wce"]), 5 * D, self.Kx, self.Kx, self.Kx, w["Wce_bf"].data_ptr())
            if self.tc_mode == 2:                    # BF16 panels of the same weights (A = W - W^T - gamma I from above)
                bp = w["Wcat_bf"].data_ptr()
                for i, src in enumerate((ph.lin_query.weight, ph.lin_key.weight, ph.lin_value.weight, w["A"])):
                    call("ctan_pack_bf16", ptr(src), D, D, D, D, bp + 2 * i * D * D)
                if self.use_tc_enc:
                    call("ctan_pack_bf16", ptr(g.enc_x.weight), D, D, D + Fn, D, w["Wenc_bf"].data_ptr())
                if self.use_tc_ro:
                    call("ctan_pack_bf16", ptr(r.lin_src.weight), Hd, D, D, D, w["Wro_bf"].data_ptr())
                    call("ctan_pack_bf16", ptr(r.lin_dst.weight), Hd, D, D, D, w["Wro_bf"].data_ptr() + 2 * Hd * D)
                if self.use_tc_ep:
                    call("ctan_pack_bf16", ptr(ph.lin_edge.weight), D, self.Fe + self.T, self.Fe + self.T, self.Kep,
                         w["Wep_bf"].data_ptr())

    def refresh_weights(self):
        """Force the weight-only operands to be re-formed at the next run() (after writing weights behind torch's back)."""
        self._w_ver = None

    def _exchange(self):
        """The one exchange of the batch, issued eagerly.  peer mode: push + wait kernels (returns the buffer parity
        used); nccl mode: torch's collective (kept OUT of CUDA-graph capture: torch's NCCL process group and stream
        capture do not mix reliably; it is still stream-ordered and needs no host sync)."""
        if self.exchange == "peer":
            xpar = self.sym.count & 1
            self._exchange_peer(xpar)
            self.sym.count += 1
            return xpar
        if self.world > 1:
            w = self.w
            exchange_rows(w["send"], w["recv"], w["rows"], self.B, self.world, self.group)
        return 0

    def _exchange_peer(self, xpar: int):
        """(capturable) store this rank's packed rows into every rank's parity-`xpar` buffer, then wait for all."""
        self.sym.push(self.w["send"], self.Q, self.q0, xpar)
        self.sym.wait(xpar)

    def _rows(self, xpar: int = 0):
        if self.exchange == "peer":
            return self.sym.rows(xpar)
        return self.w["rows"] if self.world > 1 else self.w["send"]

    def _update(self, par: int, dry: bool = False, xpar: int = 0):
        """metric accumulation + the identical ground-truth memory write-back on every rank."""
        w, m, f = self.w, self.model, self.fe[par]
        B, D = self.B, self.D
        mem, lu = m.memory.memory, m.memory.last_update
        rows = self._rows(xpar)
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):                   # the metric accumulation only feeds the final MRR: beside the
            call("ctan_accum_col", ptr(rows), B, self.row, 0, ptr(w["acc"]))          # memory write-back, not before it
        emb = ptr(rows) + 4 * (1 + self.L)
        if dry:      # warm-up: same kernels, but the state update goes to scratch tables
            sm, sl, sn, se, sd = self._scratch
            call("ctan_mem_update", ptr(sm), ptr(sl), D, ptr(sl), ptr(sl), ptr(sl), 1, emb, emb + 4 * D, self.row,
                 None, None)
            cur.wait_stream(self._side)
            return
        call("ctan_mem_update", ptr(mem), ptr(lu), D, ptr(f["b_src"]), ptr(f["b_dst"]), ptr(f["b_t"]), B, emb,
             emb + 4 * D, self.row, None, None)
        cur.wait_stream(self._side)

    def warmup(self):
        """One eager dry batch: loads every kernel and initialises the communicator before graph capture.
        Side-effect free: the state update of the dry batch is redirected to scratch tables and the cursor
        and metric accumulators are restored."""
        dev = self.dev
        self._scratch = (torch.zeros((2, self.D), device=dev), torch.zeros(2, dtype=torch.long, device=dev),
                         torch.zeros((2, self.S), dtype=torch.long, device=dev),
                         torch.full((2, self.S), -1, dtype=torch.long, device=dev),
                         torch.zeros(2, dtype=torch.int32, device=dev))
        cur = self.w["counters"].clone()
        if int(cur[0]) < self.neg_base:                      # the dry batch must lie inside the negative table
            self.w["counters"][0] = self.neg_base
        from . import _lib
        self._prep_weights()
        n0 = _lib.N_LAUNCHES
        with torch.cuda.device(dev):
            self._launch(dry=True)
        torch.cuda.synchronize(dev)
        self.launches_per_batch = _lib.N_LAUNCHES - n0      # our kernels per batch (the collective is NCCL's)
        self.w["counters"].copy_(cur)
        self.w["acc"].zero_()
        if self.use_graph:
            self._capture_all()

    def _capture_all(self):
        """Capture EVERY graph variant run() can replay (first front-end; per buffer set, with and without the
        prefetch branch) up front, so that no capture can ever land inside a caller's timed window.  Capturing
        launches nothing: the state is untouched."""
        with torch.cuda.device(self.dev):
            if self._graph is None:
                self._graphs = {}
                g0 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g0):
                    self._frontend(0)
                self._graph = g0
            xpars = (0, 1) if self.exchange == "peer" else (0,)
            for xpar in xpars:
                if self.prefetch:
                    for par in (0, 1):
                        for pf in (True, False):
                            self._graphs_for(par, pf, xpar)
                else:
                    self._graphs_for(0, False, xpar)
        torch.cuda.synchronize(self.dev)

    def _graphs_for(self, par: int, prefetch: bool, xpar: int = 0):
        """One batch as a captured graph.  1 rank / peer exchange: ONE graph (front-end | scoring | exchange kernels |
        update); NCCL exchange: two graphs around the host-issued collective."""
        key = (par, prefetch, xpar)
        if key not in self._graphs:
            ga, gb = torch.cuda.CUDAGraph(), None
            one = self.exchange != "nccl"
            with torch.cuda.graph(ga):
                if not self.prefetch:
                    self._frontend(par)                    # inline front-end: one graph per batch, no prefetch
                self._batch(par, prefetch)
                if self.exchange == "peer":
                    self._exchange_peer(xpar)
                if one:
                    self._update(par, False, xpar)
                    self._join_late()
            if not one:
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gb):
                    self._update(par)
            self._graphs[key] = (ga, gb)
        return self._graphs[key]

    def run(self, n_batches: int):
        """Evaluate the next n_batches batches.  Inside one call the front-end of batch j+1 is prefetched beside the
        scoring of batch j; the state after the call is exactly that of a batch-by-batch run."""
        if self.loader.cur_e_id + n_batches * self.B > self.n_events:
            raise ValueError("stream exhausted")
        if n_batches <= 0:
            return
        ver = (_lib.WEIGHTS_EPOCH,) + tuple((p.data_ptr(), p._version) for p in self.model.parameters())
        if ver != getattr(self, "_w_ver", None):      # weight-only operands: re-formed only when the weights changed
            self._prep_weights()
            self._w_ver = ver
        with torch.cuda.device(self.dev):
            if self.use_graph:
                if self._graph is None:
                    self._capture_all()
                if self.prefetch:
                    self._graph.replay()                   # front-end of the first batch of this call
                par = 0
                for j in range(n_batches):
                    xpar = (self.sym.count & 1) if self.sym is not None else 0
                    ga, gb = self._graphs_for(par, self.prefetch and j < n_batches - 1, xpar)
                    ga.replay()
                    if self.sym is not None:
                        self.sym.count += 1
                    if gb is not None:
                        self._exchange()
                        gb.replay()
                    if self.prefetch:
                        par ^= 1
            else:
                par = 0
                for j in range(n_batches):
                    if not self.prefetch or j == 0:
                        self._frontend(par)
                    self._batch(par, self.prefetch and j < n_batches - 1)
                    xpar = self._exchange()
                    self._update(par, False, xpar)
                    self._join_late()
                    if self.prefetch:
                        par ^= 1
        self.loader.cur_e_id += n_batches * self.B

    def scores(self) -> torch.Tensor:
        """[B, 1+n_neg] scores of the last batch (column 0 = positive), batch order."""
        rows = self._rows((self.sym.count - 1) & 1 if self.sym is not None else 0)
        return rows[: self.B, 1:1 + self.L]

    def close(self):
        """Release the symmetric peer memory (collective: every rank must call it)."""
        if self.sym is not None:
            self._graphs, self._graph = {}, None
            self.sym.close()
            self.sym = None

    def check(self):
        err = max(int(f["counts"][3].item()) for f in self.fe)
        if err:
            raise _lib.CtanLibraryError(f"neighbor lookup overflowed the engine workspace (code {err})")
