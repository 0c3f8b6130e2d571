"""Regenerate include/ctan_sm100.h from the `extern "C"` definitions in csrc/*.cu (prototype + the comment
block that precedes each definition).  Run after changing an entry point."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "non-dissipative-propagation-ctdgs_b200", "csrc")
REF = {
    "ctan_nbr_reset": "LastNeighborLoader.__init__/reset_state -- TGB/modules/neighbor_loader.py:16-23,83-85",
    "ctan_nbr_insert": "LastNeighborLoader.insert -- TGB/modules/neighbor_loader.py:41-81",
    "ctan_nbr_lookup": "LastNeighborLoader.__call__ x K + seed unique + helper[n_id]=arange -- TGB/modules/neighbor_loader.py:25-39, train_link.py:252,258-263",
    "ctan_nbr_fill": "LastNeighborLoader.__call__ (edge list part) -- TGB/modules/neighbor_loader.py:28-39",
    "ctan_mem_update": "SimpleMemory.update_state + LastAggregator -- models/memory_layers.py:132-144, TGB/modules/msg_agg.py:15-21",
    "ctan_mem_reset": "SimpleMemory.reset_state -- models/memory_layers.py:146-149",
    "ctan_edge_rel": "CTANEmbedding.forward delta-t normalisation -- models/gnn_layers.py:94-95",
    "ctan_gather_rows2": "SimpleMemory.forward + cat(z, x[n_id]) -- models/memory_layers.py:155-156, models/ctdg_models.py:444-453",
    "ctan_antisym_prep": "AntiSymmetricConv.forward: W - W^T - gamma*I -- PyG 2.3.1, call site models/gnn_layers.py:86",
    "ctan_antisym_bwd": "gradient of the above (autograd in the reference, train_link.py:279)",
    "ctan_edge_fwd": "TransformerConv.message + softmax + scatter-add, AntiSymmetricConv update, final_act -- call sites models/gnn_layers.py:82,86,98; models/ctdg_models.py:457-458",
    "ctan_edge_bwd": "gradient of the above (autograd in the reference)",
    "ctan_tanh_bwd": "gradient of final_act=tanh -- models/ctdg_models.py:457-458",
    "ctan_tenc_bwd": "gradient of TimeEncoder -- TGB/modules/time_enc.py:23-24",
    "ctan_tenc_bwd_fused": "TimeEncoder parameter gradients straight from the lin_edge output gradient (one kernel for batch-sized edge sets) -- autograd of TGB/modules/time_enc.py:23-24 + models/gnn_layers.py:97 (train_link.py:279)",
    "ctan_csr_from_sorted": "(no reference counterpart: CSR pointer of a destination-sorted edge_index)",
    "ctan_enc_fwd": "CTANEmbedding.enc_x on gathered memory rows -- models/gnn_layers.py:96",
    "ctan_eproj_fwd": "edge_attr=[msg|TimeEncoder(rel)] -> TransformerConv.lin_edge -- models/gnn_layers.py:97, TGB/modules/time_enc.py:23-24",
    "ctan_eproj_fwd_rows": "the same projection for a small edge set (a training batch): every load of a CTA in flight at once -- models/gnn_layers.py:97, TGB/modules/time_enc.py:23-24",
    "ctan_qkva_fwd": "TransformerConv lin_query/lin_key/lin_value + x @ A^T + bias -- PyG 2.3.1 (SURVEY.md App. A.5)",
    "ctan_readout_fwd": "LinkPredictor / TGBLinkPredictor / SequencePredictor.forward -- models/predictors.py:16-19,31-34,45-48",
    "ctan_readout_bwd": "gradient of the above (autograd in the reference)",
    "ctan_qkva_bwd": "gradient of the projections (autograd in the reference)",
    "ctan_qkva_wgrad": "weight / bias gradients of the q,k,v projections and of W through A = W - W^T - gamma I, accumulated straight into dW -- autograd of PyG TransformerConv / AntiSymmetricConv (train_link.py:279)",
    "ctan_enc_bwd": "gradient of enc_x (autograd in the reference)",
    "ctan_eproj_bwd": "gradient of lin_edge and of the time encoding input (autograd in the reference)",
    "ctan_stage_batch": "batch slicing, cat(src,src)/cat(dst,neg), seed cat -- train_link.py:239-256",
    "ctan_bce_loss": "BCEWithLogitsLoss(pos,1)+BCEWithLogitsLoss(neg,0) and its gradient -- train_link.py:269-270,279",
    "ctan_adam_step": "torch.optim.Adam.step -- train_link.py:281",
    "ctan_loss_generic": "criterion(pos_out, batch.y): CrossEntropyLoss / MSELoss / L1Loss (train_link.py:329,706-709) and BCEWithLogitsLoss on labels (train_sequence.py:57), with its gradient",
    "ctan_eid_mod": "e_id = e_id % data.msg.shape[0] -- train_sequence.py:41, 87",
    "ctan_zero_arena": "optimizer.zero_grad() + the zero-initialised buffers autograd accumulates into -- train_link.py:240, 279",
    "ctan_record_rows": "y_pred_confidence_list.append(...) -- per-batch scores kept for the end-of-pass metrics, train_link.py:383-395",
    "ctan_neg_sample": "NegativeSampler.sample -- negative_sampler.py:27-40",
    "ctan_sym_alloc": "(symmetric peer-memory block of the sharded evaluation; host-side, no reference counterpart)",
    "ctan_sym_free": "(host-side, no reference counterpart)",
    "ctan_ipc_export": "(CUDA IPC handle of the block; host-side, no reference counterpart)",
    "ctan_ipc_import": "(map a peer's block; host-side, no reference counterpart)",
    "ctan_ipc_close": "(host-side, no reference counterpart)",
    "ctan_xchg_push": "the exchange step of the sharded tgb_test loop (per-query scores + embeddings for the identical update at train_link.py:198): stores into every peer's buffer over NVLink",
    "ctan_xchg_wait": "consumer side of the exchange: wait until every rank's rows of the batch have arrived",
    "ctan_mrr": "TGB Evaluator._eval_hits_and_mrr (numpy branch) -- TGB/tgb/linkproppred/evaluate.py:109-120",
    "ctan_stage_eval": "per-query batch assembly of tgb_test -- train_link.py:111-142",
    "ctan_pack_eval": "per-query MRR + src/pos_dst embeddings kept for the state update -- train_link.py:150-158",
    "ctan_accum_col": "mean of the per-query metric -- train_link.py:201",
    "ctan_nbr_fill_rel": "LastNeighborLoader.__call__ edge list + CTANEmbedding delta-t -- TGB/modules/neighbor_loader.py:28-39, models/gnn_layers.py:94-95",
    "ctan_head_loss": "predictor head + BCEWithLogitsLoss + their gradients -- models/predictors.py:18-19, train_link.py:269-270,279",
    "ctan_readout_link_fused": "LinkPredictor / TGBLinkPredictor forward, BCE-with-logits loss and their gradients wrt the embeddings in one launch -- models/predictors.py:16-19, 31-34, train_link.py:269-270, 279",
    "ctan_probe": "(profiling aid, no reference counterpart)",
    "ctan_wcat_prep": "AntiSymmetricConv W - W^T - gamma*I, packed with the q/k/v weights and split for 3xTF32 -- PyG 2.3.1",
    "ctan_pack_split": "(weight packing + TF32 split for the tensor-core path; no reference counterpart)",
    "ctan_pack_split_pad": "(same, rows zero-padded to the 32-float k-block; no reference counterpart)",
    "ctan_eproj_prep": "edge_attr = cat([msg, time_enc(rel)]) as a dense TMA operand -- models/gnn_layers.py:96-97, TGB/modules/time_enc.py:23-24",
    "ctan_pack_bf16": "(BF16 weight panel for the 1e-2 tensor-core path; no reference counterpart)",
    "ctan_gemm_tc": "torch.nn.Linear on tcgen05 tensor cores (enc_x: models/gnn_layers.py:96; readout lin_src/lin_dst: models/predictors.py:16-17)",
    "ctan_enc_prep": "SimpleMemory.forward gather + node-feature columns of enc_x -- models/memory_layers.py:155-156, models/gnn_layers.py:96",
    "ctan_fuse_enc_prep": "enc_x folded into the q/k/v/x.A^T projection for frozen weights: [Wenc; Wcat Wenc] -- models/gnn_layers.py:96 + PyG App. A.5",
    "ctan_fuse_enc_prep_split": "the same panel formed every training step from the raw weights (A = W - W^T - gamma I on the fly, hi/lo split written directly) -- models/gnn_layers.py:86, 96 + PyG App. A.5",
    "ctan_enc_prep_ext": "SimpleMemory.forward gather + cat(z, x[n_id]) as a dense padded operand -- models/memory_layers.py:155-156, models/ctdg_models.py:444-453",
    "ctan_enc_qkva_tc": "enc_x and lin_query/lin_key/lin_value/x @ A^T in ONE tcgen05 contraction (evaluation, num_iters = 1) -- models/gnn_layers.py:96, PyG App. A.5",
    "ctan_head_gather": "TGBLinkPredictor.forward on pre-projected node rows -- models/predictors.py:16-19",
    "ctan_head_pack": "TGBLinkPredictor.forward on pre-projected rows + per-query reciprocal rank (TGB/tgb/linkproppred/evaluate.py:109-120) + the packed row every rank exchanges",
    "ctan_delta_t_stats_ws_bytes": "(workspace size of ctan_delta_t_stats; host-side query)",
    "ctan_delta_t_stats": "compute_stats: mean/std of per-node inter-event times -- utils.py:46-92",
    "ctan_set_pdl": "(host-side switch for programmatic dependent launch; no reference counterpart)",
    "ctan_hub_list": "(work list of high-degree destination rows for ctan_edge_fwd; no reference counterpart)",
    "ctan_dx_tc": "gradient of the projections w.r.t. the node state (autograd in the reference) on tcgen05",
    "ctan_qkva_fwd_tc": "TransformerConv lin_query/lin_key/lin_value + x @ A^T + bias on tcgen05 tensor cores -- PyG 2.3.1 (SURVEY.md App. A.5)",
}
out = ['''/* ctan_sm100.h -- C ABI of libctan_sm100.so (hand-written CUDA for sm_100a, NVIDIA B200).
 *
 * B200-native replacement of the per-event-batch hot path of CTAN
 * (gravins/non-dissipative-propagation-CTDGs).  The reference has no FFI: its boundary is the Python module
 * API (models/ctdg_models.py, models/gnn_layers.py, models/memory_layers.py, models/predictors.py and PyG's
 * LastNeighborLoader).  Each entry point below names the reference code it replaces; the Python mirror of the
 * reference API (non-dissipative-propagation-ctdgs_b200/*.py) binds these symbols with ctypes.
 *
 * Conventions: plain pointers and sizes only (no torch types).  All pointers are DEVICE pointers unless
 * named *_host.  ids / event ids / times are int64, features fp32, CSR pointers int32.  Every function
 * enqueues its kernels on `stream` and returns 0, a cudaError_t (>0), or a negative CTAN_ERR_* code; none
 * allocates, synchronises or keeps global state.  An `int n, const int* n_dev` pair means: when n_dev is
 * non-NULL the true size is read on the device (produced by an earlier kernel) and n is only its upper
 * bound.  Generated by scripts/gen_header.py from csrc/*.cu -- do not edit by hand. */
#ifndef CTAN_SM100_H
#define CTAN_SM100_H
#include <stdint.h>
#include <cuda_runtime_api.h>

#define CTAN_ERR_BAD_ARG (-1)
#define CTAN_ERR_UNSUPPORTED (-2)

#ifdef __cplusplus
extern "C" {
#endif
''']
names = []
for f in ("loader.cu", "edge.cu", "dense.cu", "step.cu", "tc_gemm.cu", "stats.cu", "sampler.cu", "sym.cu"):
    src = open(os.path.join(CSRC, f)).read()
    out.append(f"/* ---------------------------------------------------------------- {f} */")
    for m in re.finditer(r'((?:^//[^\n]*\n)*)extern "C" int (\w+)\(([^{]*?)\)\s*\{', src, re.M):
        comment, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        if name == "ctan_tc_trace_read":          # exists only in -DCTAN_TC_TRACE profiling builds
            continue
        names.append(name)
        out.append(f"/* replaces: {REF.get(name, '(internal helper)')} */")
        if comment.strip():
            out.append(comment.rstrip())
        out.append(f"int {name}({args});\n")
out.append('''#ifdef __cplusplus
}
#endif
#endif /* CTAN_SM100_H */''')
os.makedirs(os.path.join(ROOT, "include"), exist_ok=True)
open(os.path.join(ROOT, "include", "ctan_sm100.h"), "w").write("\n".join(out) + "\n")
print(len(names), "entry points")
