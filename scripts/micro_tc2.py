"""tcgen05 GEMM cost model: one tile per SM (M = 148*128, Nout = 128), sweep K -> per-k-block slope and
fixed per-tile cost, for 3xTF32 (mode 0) and TF32 (mode 1); then the eval shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ctan_b200
from ctan_b200._lib import call, ptr
dev = torch.device("cuda:0")


def timeit(fn, R=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(R):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * R) * 1e3


def run(M, K, Nout, mode, dev_sized=False):
    A = torch.randn(M, K, device=dev)
    W = torch.randn(Nout, K, device=dev) / K ** 0.5
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    call("ctan_pack_split", ptr(W), Nout, K, K, ptr(hi), ptr(lo))
    b = torch.randn(Nout, device=dev)
    C = torch.empty(M, Nout, device=dev)
    Md = torch.tensor([M], dtype=torch.int32, device=dev)
    us = timeit(lambda: call("ctan_gemm_tc", ptr(A), M, ptr(Md) if dev_sized else None, K, ptr(hi), ptr(lo), Nout,
                             ptr(b), None, None, None, Nout, None, 0, ptr(C), mode))
    return us


for mode in (0, 1):
    for K in (32, 256, 1024, 2048):
        us = run(148 * 128, K, 128, mode)
        print(f"one tile/SM  mode={mode} K={K:5d} ({K//32:3d} k-blocks): {us:8.2f} us")
for M, K, Nout in [(6027, 256, 256), (6027, 256, 512), (6027, 256, 1024), (40000, 256, 1024)]:
    for mode in (0, 1):
        for ds in (True,):
            us = run(M, K, Nout, mode, ds)
            print(f"M={M} K={K} Nout={Nout} mode={mode} dev_sized={ds}: {us:8.2f} us  "
                  f"{2.0 * M * K * Nout / us / 1e6:7.1f} useful TFLOP/s")
