"""torch.autograd.Function wrappers around the C-ABI kernels (forward AND hand-written backward).

`ctan_embed`   = CTANEmbedding.forward  (models/gnn_layers.py:93-98 + PyG AntiSymmetricConv/TransformerConv)
`link_readout` = LinkPredictor / TGBLinkPredictor / SequencePredictor.forward (models/predictors.py)

PyTorch is plumbing here (allocation, autograd bookkeeping, stream); every arithmetic op is one of
our kernels.  No fallback path exists."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr


def _f32(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


class _EmbedFn(torch.autograd.Function):
    """z = [final_act](AntiSymmetricConv_K(enc_x(z0), edges)).  Inputs with gradients are the 15
    parameter tensors; z0 / tables / indices are constants (the memory is stored detached,
    models/memory_layers.py:144)."""

    @staticmethod
    def forward(ctx, z0, lu, n_id, esrc, rowptr, msg, t_e, hp,
                Wenc, benc, Wq, bq, Wk, bk, Wv, bv, We, W, b, tw, tb):
        K, eps, gamma, mean, std, final_tanh = hp
        N, Kin = z0.shape
        D = W.size(0)
        E = esrc.numel()
        Fe = msg.size(1) if msg.dim() == 2 else 0
        T = tb.numel()
        dev = z0.device
        need_grad = any(ctx.needs_input_grad[8:])
        with torch.cuda.device(dev):
            rel = torch.empty(E, device=dev)
            call("ctan_edge_rel", ptr(lu, torch.long), ptr(n_id, torch.long) if n_id is not None else None,
                 ptr(esrc, torch.long), ptr(t_e, torch.long), None, E, None, float(mean), float(std), ptr(rel))
            nk = K if need_grad else min(K, 1)
            X = torch.empty((K + 1 if need_grad else 2, N, D), device=dev)
            call("ctan_enc_fwd", None, D, None, Kin - D, None, ptr(z0), N, None, ptr(Wenc), ptr(benc), ptr(X[0]))
            EP = torch.empty((E, D), device=dev)
            call("ctan_eproj_fwd", ptr(msg), Fe, None, ptr(rel), ptr(tw), ptr(tb), T, ptr(We), E, None, D, ptr(EP))
            A = torch.empty((D, D), device=dev)
            use_tc = _lib.use_tensor_cores(D)
            if use_tc:
                Whi, Wlo = torch.empty((4 * D, D), device=dev), torch.empty((4 * D, D), device=dev)
                call("ctan_wcat_prep", ptr(Wq), ptr(Wk), ptr(Wv), ptr(W), float(gamma), D, ptr(Whi), ptr(Wlo), ptr(A), None, None)
            else:
                call("ctan_antisym_prep", ptr(W), float(gamma), D, ptr(A))
            QKVA = torch.empty((max(nk, 1), N, 4 * D), device=dev)
            TH = torch.empty((K, N, D), device=dev) if need_grad else None
            ALPHA = torch.empty((K, max(E, 1)), device=dev) if need_grad else None
            Z = torch.empty((N, D), device=dev) if final_tanh else None
            for k in range(K):
                xi = X[k] if need_grad else X[k & 1]
                xo = X[k + 1] if need_grad else X[(k + 1) & 1]
                qk = QKVA[k] if need_grad else QKVA[0]
                if use_tc:
                    call("ctan_qkva_fwd_tc", ptr(xi), N, None, D, ptr(Whi), ptr(Wlo), ptr(bq), ptr(bk), ptr(bv),
                         ptr(b), ptr(qk), _lib.TC_MODE)
                else:
                    call("ctan_qkva_fwd", ptr(xi), N, None, D, ptr(Wq), ptr(Wk), ptr(Wv), ptr(A), ptr(bq), ptr(bk),
                         ptr(bv), ptr(b), ptr(qk))
                last = k == K - 1
                call("ctan_edge_fwd", ptr(qk), ptr(EP), ptr(rowptr, torch.int32), ptr(esrc), N, None, D, ptr(xi),
                     float(eps), ptr(xo), ptr(TH[k]) if need_grad else None, ptr(ALPHA[k]) if need_grad else None,
                     ptr(Z) if (last and final_tanh) else None, None, 0, 0)
            out = Z if final_tanh else (X[K] if need_grad else X[K & 1])
            if K == 0:
                raise ValueError("num_iters must be >= 1")
        if need_grad:
            ctx.hp = hp
            ctx.save_for_backward(z0, esrc, rowptr, msg, rel, X, EP, A, QKVA, TH, ALPHA, out,
                                  Wq, Wk, Wv, We, tw, tb)
        return out

    @staticmethod
    def backward(ctx, dz):
        K, eps, gamma, mean, std, final_tanh = ctx.hp
        (z0, esrc, rowptr, msg, rel, X, EP, A, QKVA, TH, ALPHA, out, Wq, Wk, Wv, We, tw, tb) = ctx.saved_tensors
        N, Kin = z0.shape
        D = A.size(0)
        E = esrc.numel()
        Fe = msg.size(1) if msg.dim() == 2 else 0
        T = tb.numel()
        dev = z0.device
        with torch.cuda.device(dev):
            g = _f32(dz)
            if final_tanh:
                g2 = torch.empty_like(g)
                call("ctan_tanh_bwd", ptr(g), ptr(out), N, None, D, ptr(g2))
                g = g2
            flat = torch.zeros(4 * D * D + 4 * D + D * Kin + D + D * (Fe + T) + 2 * T, device=dev)
            o = 0

            def take(n, shape):
                nonlocal o
                v = flat[o:o + n].view(shape)
                o += n
                return v
            dWq, dWk, dWv, dA = (take(D * D, (D, D)) for _ in range(4))
            dbq, dbk, dbv, db = (take(D, (D,)) for _ in range(4))
            dWenc, dbenc = take(D * Kin, (D, Kin)), take(D, (D,))
            dWe = take(D * (Fe + T), (D, Fe + T))
            dtw, dtb = take(T, (T,)), take(T, (T,))
            dEP = torch.empty((E, D), device=dev)
            DS = torch.empty(max(E, 1), device=dev)
            DQKVA = torch.empty((N, 4 * D), device=dev)
            for k in range(K - 1, -1, -1):
                DQKVA.zero_()
                call("ctan_edge_bwd", ptr(QKVA[k]), ptr(EP), ptr(rowptr), ptr(esrc), N, None, D, ptr(g), ptr(TH[k]),
                     ptr(ALPHA[k]), float(eps), ptr(DQKVA), ptr(dEP), int(k != K - 1), ptr(DS))
                gn = torch.zeros((N, D), device=dev)        # the dX contraction splits K and accumulates
                call("ctan_qkva_bwd", ptr(DQKVA), ptr(X[k]), ptr(g), N, None, D, ptr(Wq), ptr(Wk), ptr(Wv), ptr(A),
                     ptr(gn), ptr(dWq), ptr(dWk), ptr(dWv), ptr(dA), ptr(dbq), ptr(dbk), ptr(dbv), ptr(db), 0)
                g = gn
            call("ctan_enc_bwd", ptr(g), ptr(z0), N, None, D, Kin - D, ptr(dWenc), ptr(dbenc), 0)
            if E > 0:
                dTenc = torch.empty((E, T), device=dev)
                call("ctan_eproj_bwd", ptr(dEP), ptr(msg), Fe, None, ptr(rel), ptr(tw), ptr(tb), T, ptr(We), E, None,
                     D, ptr(dWe), ptr(dTenc), 0)
                call("ctan_tenc_bwd", ptr(dTenc), ptr(rel), ptr(tw), ptr(tb), E, None, T, ptr(dtw), ptr(dtb))
            dW = torch.empty((D, D), device=dev)
            call("ctan_antisym_bwd", ptr(dA), D, ptr(dW))
        return (None,) * 8 + (dWenc, dbenc, dWq, dbq, dWk, dbk, dWv, dbv, dWe, dW, db, dtw, dtb)


def ctan_embed(z0, last_update, n_id, edge_src, rowptr, msg, t_e, *, num_iters, epsilon, gamma, mean_delta_t,
               std_delta_t, final_tanh, Wenc, benc, Wq, bq, Wk, bk, Wv, bv, We, W, b, tw, tb):
    """z0 [N, D+Fn] fp32 (memory rows | node features); last_update: the [num_nodes] table when n_id is
    given, else the per-local-node [N] array; edge_src [E] local source ids of a destination-sorted edge
    list with CSR `rowptr` [N+1] int32; msg [E,Fe]; t_e [E] int64.  Returns z [N,D]."""
    hp = (int(num_iters), float(epsilon), float(gamma), float(mean_delta_t), float(std_delta_t), bool(final_tanh))
    return _EmbedFn.apply(_f32(z0), last_update.contiguous(), n_id, edge_src.contiguous(), rowptr.contiguous(),
                          _f32(msg), t_e.contiguous(), hp, Wenc, benc, Wq, bq, Wk, bk, Wv, bv, We, W, b,
                          tw.view(-1), tb)


class _ReadoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, src, dst, id_map, Wsrc, bsrc, Wdst, bdst, Wfin, bfin):
        P = dst.numel()
        D = z.size(1)
        Hd, O = Wdst.size(0), Wfin.size(0)
        dev = z.device
        with torch.cuda.device(dev):
            # training: split-K (atomic, needs a zeroed H); no-grad scoring: plain stores, bit-reproducible
            split = any(ctx.needs_input_grad)
            H = torch.zeros((P, Hd), device=dev) if split else torch.empty((P, Hd), device=dev)
            out = torch.empty((P, O), device=dev)
            call("ctan_readout_fwd", ptr(z), D, ptr(src, torch.long) if src is not None else ptr(dst, torch.long),
                 ptr(dst, torch.long), ptr(id_map, torch.long) if id_map is not None else None, P, None, Hd, O,
                 ptr(Wsrc) if Wsrc is not None else None, ptr(bsrc) if bsrc is not None else None, ptr(Wdst),
                 ptr(bdst), ptr(Wfin), ptr(bfin), ptr(H), ptr(out), int(split))
        # the caller's id map (`assoc`) is a scratch table it overwrites on the next batch, while backward may
        # run many batches later (train_sequence.py:30-65): keep the resolved rows, not the table
        if id_map is not None:
            src = id_map[src] if src is not None else None
            dst = id_map[dst]
        ctx.save_for_backward(z, src, dst, H, Wsrc, Wdst, Wfin)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, src, dst, H, Wsrc, Wdst, Wfin = ctx.saved_tensors
        id_map = None
        P = dst.numel()
        N, D = z.shape
        Hd, O = Wdst.size(0), Wfin.size(0)
        dev = z.device
        has_src = Wsrc is not None
        with torch.cuda.device(dev):
            dout = _f32(dout)
            dz = torch.zeros((N, D), device=dev)
            flat = torch.zeros(2 * Hd * D + 2 * Hd + O * Hd + O, device=dev)
            dWsrc, dWdst = flat[:Hd * D].view(Hd, D), flat[Hd * D:2 * Hd * D].view(Hd, D)
            o = 2 * Hd * D
            dbsrc, dbdst = flat[o:o + Hd], flat[o + Hd:o + 2 * Hd]
            o += 2 * Hd
            dWfin, dbfin = flat[o:o + O * Hd].view(O, Hd), flat[o + O * Hd:]
            dHid = torch.empty((P, Hd), device=dev)
            call("ctan_readout_bwd", ptr(dout), ptr(H), ptr(z), D, ptr(src) if src is not None else ptr(dst), ptr(dst),
                 ptr(id_map) if id_map is not None else None, P, None, Hd, O, ptr(Wsrc) if has_src else None,
                 ptr(Wdst), ptr(Wfin), ptr(dHid), ptr(dz), ptr(dWsrc) if has_src else None,
                 ptr(dbsrc) if has_src else None, ptr(dWdst), ptr(dbdst), ptr(dWfin), ptr(dbfin), 0)
        return (dz, None, None, None, dWsrc if has_src else None, dbsrc if has_src else None, dWdst, dbdst,
                dWfin, dbfin)


def link_readout(z, src, dst, id_map, Wsrc, bsrc, Wdst, bdst, Wfin, bfin):
    """out[p] = Wfin relu(Wsrc z[row(src_p)] + bsrc + Wdst z[row(dst_p)] + bdst) + bfin with
    row(g) = id_map[g] (or g when id_map is None).  Wsrc=None gives the destination-only predictor."""
    return _ReadoutFn.apply(_f32(z), src.contiguous() if src is not None else None, dst.contiguous(),
                            id_map, Wsrc, bsrc, Wdst, bdst, Wfin, bfin)
