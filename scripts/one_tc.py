"""A few launches of the tcgen05 GEMM at one shape (for ncu --set full)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ctan_b200
from ctan_b200._lib import call, ptr
M, K, Nout, mode = (int(v) for v in (sys.argv[1:5] + ["40000", "256", "1024", "0"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
A = torch.randn(M, K, device=dev)
W = torch.randn(Nout, K, device=dev) / K ** 0.5
hi, lo = torch.empty_like(W), torch.empty_like(W)
call("ctan_pack_split", ptr(W), Nout, K, K, ptr(hi), ptr(lo))
b = torch.randn(Nout, device=dev)
C = torch.empty(M, Nout, device=dev)
for _ in range(5):
    call("ctan_gemm_tc", ptr(A), M, None, K, ptr(hi), ptr(lo), Nout, ptr(b), None, None, None, Nout, None, 0, ptr(C), mode)
torch.cuda.synchronize()
ref = A.double() @ W.double().t() + b.double()
print("max rel err", float((C.double() - ref).abs().max() / ref.abs().max()))
