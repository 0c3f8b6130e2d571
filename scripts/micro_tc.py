"""qkvA projection: tcgen05 (3xTF32) vs CUDA-core fp32, warm, 50 launches per graph replay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ctan_b200
from ctan_b200._lib import call, ptr
dev = torch.device("cuda:0")
for N, D in [(700, 128), (3000, 128), (6000, 256), (40000, 256)]:
    X = torch.randn(N, D, device=dev)
    W = [torch.randn(D, D, device=dev) / D ** 0.5 for _ in range(4)]
    b = [torch.randn(D, device=dev) for _ in range(4)]
    hi, lo, A = torch.empty(4 * D, D, device=dev), torch.empty(4 * D, D, device=dev), torch.empty(D, D, device=dev)
    call("ctan_wcat_prep", ptr(W[0]), ptr(W[1]), ptr(W[2]), ptr(W[3]), 0.1, D, ptr(hi), ptr(lo), ptr(A), None, None)
    out = torch.empty(N, 4 * D, device=dev)
    fns = {"tc3x": lambda: call("ctan_qkva_fwd_tc", ptr(X), N, None, D, ptr(hi), ptr(lo), ptr(b[0]), ptr(b[1]), ptr(b[2]), ptr(b[3]), ptr(out), 0),
           "tc1x": lambda: call("ctan_qkva_fwd_tc", ptr(X), N, None, D, ptr(hi), ptr(lo), ptr(b[0]), ptr(b[1]), ptr(b[2]), ptr(b[3]), ptr(out), 1),
           "simt": lambda: call("ctan_qkva_fwd", ptr(X), N, None, D, ptr(W[0]), ptr(W[1]), ptr(W[2]), ptr(A), ptr(b[0]), ptr(b[1]), ptr(b[2]), ptr(b[3]), ptr(out))}
    for name, fn in fns.items():
        fn(); torch.cuda.synchronize()
        R = 20
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
        us = e0.elapsed_time(e1) / (5 * R) * 1e3
        fl = 2.0 * N * D * 4 * D * (3 if name.startswith("tc") else 1)
        print(f"N={N:6d} D={D:4d} {name}: {us:9.2f} us/call   {2.0*N*D*4*D/us/1e6:8.2f} useful TFLOP/s")
