"""`compute_stats` on the device.  Mirrors utils.py:46-92.

    compute_stats(src, dst, t, init_time)                               the training split's events (CUDA tensors)
    compute_stats_from_data(data, split, init_time, sequence_prediction) the reference's own signature (main.py:302,
                                                                        main_tgb_ctan.py:270): splits `data`, concatenates
                                                                        the sequences in the sequence case

Both return (mean_delta_t, std_delta_t) as Python floats (fp64 accumulation on the device)."""
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


def compute_stats_from_data(data, split, init_time, sequence_prediction: bool = False, device=None):
    """utils.compute_stats(data, split, init_time, sequence_prediction) (utils.py:46-92) with the Python loop over all
    training events replaced by the device kernel.  `data` needs `train_val_test_split(val_ratio, test_ratio)` like
    PyG's TemporalData (or, for sequences, the dataset object of datasets/label_prediction_sequence.py)."""
    train_data, _, _ = data.train_val_test_split(val_ratio=split[0], test_ratio=split[1])
    if sequence_prediction:
        src = torch.cat([d.src for d in train_data])
        dst = torch.cat([d.dst for d in train_data])
        t = torch.cat([d.t for d in train_data])
    else:
        src, dst, t = train_data.src, train_data.dst, train_data.t
    dev = torch.device(device if device is not None else (src.device if src.is_cuda else "cuda"))
    return compute_stats(src.to(dev), dst.to(dev), t.to(dev), int(init_time))
