"""ctypes binding of the C-ABI CUDA library `csrc/libctan_sm100.so` (declared in
`include/ctan_sm100.h`).  There is NO fallback: if the library is missing or a call fails, we raise."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libctan_sm100.so")

P, I, L, F = c_void_p, c_int, c_int64, c_float

# name -> argument types (the trailing cudaStream_t is appended automatically)
SIGNATURES = {
    # loader.cu
    "ctan_nbr_reset": [P, P, P, L, I],
    "ctan_nbr_insert": [P, P, P, P, P, I, L, P, I, P],
    "ctan_nbr_lookup": [P, P, L, I, P, I, P, I, P, P, P, P, I, P, I, P, P, I, P],
    "ctan_nbr_fill": [P, P, I, P, P, P, I, P, P, P, P, P, P],
    "ctan_nbr_fill_rel": [P, P, I, P, P, P, I, P, P, P, P, P, P, P, P, F, F, P],
    "ctan_mem_update": [P, P, I, P, P, P, I, P, P, I, P, P],
    "ctan_mem_reset": [P, P, L, I, L],
    # edge.cu
    "ctan_edge_rel": [P, P, P, P, P, I, P, F, F, P],
    "ctan_gather_rows2": [P, I, P, I, P, I, P, P],
    "ctan_antisym_prep": [P, F, I, P],
    "ctan_antisym_bwd": [P, I, P],
    "ctan_edge_fwd": [P, P, P, P, I, P, I, P, F, P, P, P, P, P, I, I],
    "ctan_hub_list": [P, I, P, P, I],
    "ctan_edge_bwd": [P, P, P, P, I, P, I, P, P, P, F, P, P, I, P],
    "ctan_tanh_bwd": [P, P, I, P, I, P],
    "ctan_tenc_bwd": [P, P, P, P, I, P, I, P, P],
    "ctan_tenc_bwd_fused": [P, P, I, I, I, P, P, P, I, P, P, P],
    "ctan_csr_from_sorted": [P, I, I, P, P],
    # dense.cu
    "ctan_enc_fwd": [P, I, P, I, P, P, I, P, P, P, P],
    "ctan_eproj_fwd": [P, I, P, P, P, P, I, P, I, P, I, P],
    "ctan_eproj_fwd_rows": [P, I, P, P, P, P, I, P, I, P, I, P],
    "ctan_qkva_fwd": [P, I, P, I, P, P, P, P, P, P, P, P, P],
    "ctan_readout_fwd": [P, I, P, P, P, I, P, I, I, P, P, P, P, P, P, P, P, I],
    "ctan_readout_bwd": [P, P, P, I, P, P, P, I, P, I, I, P, P, P, P, P, P, P, P, P, P, P, I],
    "ctan_head_loss": [P, I, I, P, P, I, P, P, P, P, I, P],
    "ctan_readout_link_fused": [P, I, P, P, P, I, I, I, P, P, P, P, P, P, P, P, P, P, P, P, I, P],
    "ctan_qkva_bwd": [P, P, P, I, P, I, P, P, P, P, P, P, P, P, P, P, P, P, P, I],
    "ctan_qkva_wgrad": [P, P, I, P, I, P, P, P, P, P, P, P, P, I],
    "ctan_enc_bwd": [P, P, I, P, I, I, P, P, I],
    "ctan_eproj_bwd": [P, P, I, P, P, P, P, I, P, I, P, I, P, P, I],
}

_lib = None


class CtanLibraryError(RuntimeError):
    pass


def lib():
    """Load (once) and return the CDLL.  Raises loudly when the CUDA library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CtanLibraryError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(ctan_b200 has no CPU / PyTorch fallback by design)")
    dll = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(dll, name)                    # AttributeError if the symbol is missing
        fn.argtypes = list(args) + [c_void_p]
        fn.restype = c_int
    _optional_signatures(dll)
    _lib = dll
    return dll


OPTIONAL_SIGNATURES = {}

SIGNATURES.update({
    # step.cu
    "ctan_stage_batch": [P, P, P, P, I, I, P, I, P, P, P, P, P, P, P, I, P],
    "ctan_bce_loss": [P, I, I, P, P, I, P],
    "ctan_adam_step": [P, P, P, P, I, F, F, F, F, F, L, P, P],
    "ctan_loss_generic": [P, I, I, I, P, F, P, P, P, I, P],
    "ctan_eid_mod": [P, I, P, I, P],
    "ctan_zero_arena": [P, L, I, P, P, P, P, P, P, P, P, P, P, I],
    "ctan_record_rows": [P, I, P, I, P],
    # sym.cu (the host-side ctan_sym_* / ctan_ipc_* setup calls take no stream: bound in sym.py)
    "ctan_xchg_push": [P, I, I, I, I, P, I, I, I, ctypes.c_size_t, P, P],
    "ctan_xchg_wait": [P, I, P, P],
    # sampler.cu
    "ctan_neg_sample": [P, L, P, L, P, I, ctypes.c_uint64, ctypes.c_uint64, I, P, P],
    "ctan_mrr": [P, P, I, I, P],
    "ctan_stage_eval": [P, P, P, P, I, I, I, I, P, P, P, P, P, P, P],
    "ctan_pack_eval": [P, P, P, P, P, I, I, I, P],
    "ctan_accum_col": [P, I, I, I, P],
    "ctan_probe": [P, I],
    # tc_gemm.cu
    "ctan_wcat_prep": [P, P, P, P, F, I, P, P, P, P, P],
    "ctan_dx_tc": [P, P, I, P, I, P, P, P, I, I],
    "ctan_pack_split": [P, I, I, I, P, P],
    "ctan_pack_split_pad": [P, I, I, I, I, P, P],
    "ctan_pack_bf16": [P, I, I, I, I, P],
    "ctan_eproj_prep": [P, I, P, P, P, P, I, I, P, I, P],
    "ctan_gemm_tc": [P, I, P, I, P, P, I, P, P, P, P, I, P, I, P, I],
    "ctan_enc_prep": [P, I, P, I, P, I, P, P, P, P],
    "ctan_head_gather": [P, I, P, P, P, I, P, P, P],
    "ctan_fuse_enc_prep": [P, P, P, P, P, P, P, P, P, P, I, I, I, P, P],
    "ctan_fuse_enc_prep_split": [P, P, P, P, F, P, P, P, P, P, P, I, I, I, P, P, P],
    "ctan_enc_prep_ext": [P, I, P, I, P, I, P, I, P],
    "ctan_enc_qkva_tc": [P, I, P, I, P, P, I, P, P, P, I],
    "ctan_head_pack": [P, I, P, I, P, P, P, I, I, P, P, P],
    "ctan_set_pdl": [I],
    "ctan_delta_t_stats_ws_bytes": [L, P],
    "ctan_delta_t_stats": [P, P, P, L, L, P, L, P],
    "ctan_qkva_fwd_tc": [P, I, P, I, P, P, P, P, P, P, P, I],
})


def _optional_signatures(dll):
    for name, args in OPTIONAL_SIGNATURES.items():
        fn = getattr(dll, name)
        fn.argtypes = list(args) + [c_void_p]
        fn.restype = c_int


HOST_ONLY = ["ctan_sym_alloc", "ctan_sym_free", "ctan_ipc_export", "ctan_ipc_import", "ctan_ipc_close"]


def exported_symbols():
    return list(SIGNATURES) + list(OPTIONAL_SIGNATURES) + HOST_ONLY


def use_tensor_cores(D: int) -> bool:
    """The tcgen05/TMA projection needs D % 32 == 0 (128-byte K blocks, 128-wide N tiles).  CTAN_NO_TC=1
    forces the CUDA-core contraction (A/B comparisons)."""
    return D % 32 == 0 and os.environ.get("CTAN_NO_TC", "0") != "1"


# precision of the tensor-core projection: 0 = 3xTF32 (fp32 parity, 1e-4 bar), 1 = single-pass TF32 (1e-2 bar)
TC_MODE = int(os.environ.get("CTAN_TC_MODE", "0"))


def stream_ptr() -> int:
    # the raw-stream query is ~30x cheaper than torch.cuda.current_stream().cuda_stream (which builds a Stream
    # object per call); the strict-API path makes ~20 C-ABI calls per step
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CtanLibraryError("ctan_b200 kernels need CUDA tensors (no CPU path)")
    if not t.is_contiguous():
        raise CtanLibraryError("ctan_b200 kernels need contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise CtanLibraryError(f"expected dtype {dtype}, got {t.dtype}")
    return t.data_ptr()


# kernels launched by each C entry point (for launch accounting; ctan_nbr_lookup adds one per hop)
LAUNCHES = {"ctan_set_pdl": 0, "ctan_delta_t_stats_ws_bytes": 0, "ctan_delta_t_stats": 9, "ctan_nbr_insert": 2, "ctan_nbr_lookup": 1, "ctan_readout_fwd": 2, "ctan_readout_bwd": 2,
            "ctan_qkva_bwd": 2, "ctan_eproj_bwd": 0, "ctan_sym_alloc": 0}

# Bumped by every writer of model weights that bypasses torch's tensor versioning (StepEngine's Adam kernel updates
# the flat parameter buffer through raw pointers); readers of derived weight operands compare it with their last prep.
WEIGHTS_EPOCH = 0

# Optional instrumentation (bench.py): TRACE = list -> every call appends (name, start_event, end_event)
TRACE = None
N_LAUNCHES = 0


# Optional timeline probing (scripts/timeline_step.py): PROBE = (ts_tensor, names_list) -> after every call made
# on the stream that was current when probing was switched on, a probe kernel stores %globaltimer.
PROBE = None
PROBE_SIDE = False


def call(name: str, *args):
    global N_LAUNCHES
    fn = getattr(lib(), name)
    N_LAUNCHES += LAUNCHES.get(name, 1) + (args[7] if name == "ctan_nbr_lookup" else 0)
    if name == "ctan_eproj_bwd":                            # one contraction per non-NULL output (dWe, dTenc)
        N_LAUNCHES += (args[12] is not None) + (args[13] is not None)
    if name == "ctan_edge_fwd" and args[13] is not None and args[15] == 0:
        N_LAUNCHES += 1                                     # the hub-row kernel (modes 1 / 2 launch one kernel each)
    if PROBE is not None and name != "ctan_probe":
        ts, names, main = PROBE
        rc = fn(*args, stream_ptr())
        if rc != 0:
            raise CtanLibraryError(f"{name} failed with code {rc}")
        cur = torch.cuda.current_stream().cuda_stream
        if cur == main:
            lib().ctan_probe(ts.data_ptr(), len(names), stream_ptr())
            names.append(name)
        elif PROBE_SIDE and len(names) < ts.numel():         # side-stream calls too (scripts/timeline_step.py ALL=1)
            lib().ctan_probe(ts.data_ptr(), len(names), stream_ptr())
            names.append(f"{name} [side {cur & 0xFFFF:04x}]")
        return
    if TRACE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args, stream_ptr())
        e1.record()
        TRACE.append((name, e0, e1))
    else:
        rc = fn(*args, stream_ptr())
    if rc != 0:
        raise CtanLibraryError(f"{name} failed with code {rc}"
                               + (f" ({_cuda_err(rc)})" if rc > 0 else " (bad argument / unsupported size)"))


def _cuda_err(rc: int) -> str:
    try:
        from cuda.bindings import runtime as cudart  # type: ignore
        return str(cudart.cudaGetErrorString(cudart.cudaError_t(rc))[1])
    except Exception:
        return "cudaError"
