"""tcgen05 contraction, single-CTA form vs two-CTA clusters with TMA-multicast weight tiles (`mode` bit 6), at the
evaluation shapes (tgbl-comment-shaped batch) and at the saturating shape of profiles/r01_ncu_tc_gemm.txt."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ctan_b200._lib import call, ptr

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (M, K, Nout, what) in ((6093, 288, 1280, "eval: folded enc_x + projection"), (6760, 288, 256, "eval: lin_edge"),
                           (6093, 256, 512, "eval: readout node projection"), (40000, 256, 1024, "saturating shape"),
                           (1006, 288, 1280, "8-rank shard"), (640, 160, 640, "training batch")):
    A = torch.randn(M + 128, K, generator=g).to(dev)
    W = (torch.randn(Nout, K, generator=g) / K ** 0.5).to(dev)
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    call("ctan_pack_split", ptr(W), Nout, K, K, ptr(hi), ptr(lo))
    b = torch.randn(Nout, device=dev)
    C = torch.empty(M, Nout, device=dev)
    res = {}
    for mode in (0, 0x40):
        fn = lambda: call("ctan_gemm_tc", ptr(A), M, None, K, ptr(hi), ptr(lo), Nout, ptr(b), None, None, None, Nout, None, 0,
                          ptr(C), mode)
        res[mode] = timeit(fn)
    fl = 2.0 * M * K * Nout
    print(f"{what:34s} [{M},{K}]x[{K},{Nout}]: single {res[0]:7.1f} us ({fl / res[0] / 1e6:6.1f} useful TFLOP/s)   "
          f"pair {res[0x40]:7.1f} us ({fl / res[0x40] / 1e6:6.1f})", flush=True)
