"""`compute_stats` on the device.  Mirrors utils.py:46-92: the caller passes the training
split's (src, dst, t); returns (mean_delta_t, std_delta_t) as Python floats (fp64 on the device)."""
from __future__ import annotations

import ctypes

import torch

from ._lib import CtanLibraryError, call, lib, ptr


def compute_stats(src: torch.Tensor, dst: torch.Tensor, t: torch.Tensor, init_time: int = 0):
    if not (src.is_cuda and dst.is_cuda and t.is_cuda):
        raise CtanLibraryError("compute_stats needs CUDA tensors (no CPU path)")
    n = src.numel()
    if n == 0:
        raise ValueError("compute_stats of an empty split")
    src, dst, t = (v.contiguous().to(torch.long) for v in (src, dst, t))
    need = ctypes.c_uint64(0)
    call("ctan_delta_t_stats_ws_bytes", n, ctypes.byref(need))
    with torch.cuda.device(src.device):
        ws = torch.empty(int(need.value) + 256, dtype=torch.uint8, device=src.device)
        base = ws.data_ptr()
        off = (-base) % 256
        stats = torch.zeros(2, dtype=torch.float64, device=src.device)
        call("ctan_delta_t_stats", ptr(src), ptr(dst), ptr(t), n, int(init_time), base + off, int(need.value), ptr(stats))
        m, s = stats.tolist()
    return m, s
