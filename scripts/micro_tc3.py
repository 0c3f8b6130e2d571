"""Phase timeline of the tcgen05 contraction at the training-batch shapes (profiling build with -DCTAN_TC_TRACE).

    python scripts/micro_tc3.py          # builds gpurun_out/libctan_trace.so, runs qkvA forward and dX at N = 640, D = 128

Prints, per CTA that did work, the offsets (us, from the kernel's first CTA entry) of: entry, set-up done, first TMA
issued, each k-block's arrival in shared memory, first MMA, last commit, accumulator complete (epilogue start), epilogue
end, kernel end -- and the CUDA-event duration of the warm launch."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

CSRC = os.path.join(ROOT, "non-dissipative-propagation-ctdgs_b200", "csrc")
OUT = os.path.join(ROOT, "gpurun_out", "libctan_trace.so")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
srcs = [os.path.join(CSRC, f) for f in ("tc_gemm.cu", "step.cu")]
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-DCTAN_TC_TRACE",
                       "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", OUT, *srcs, "-lcuda"])
lib = ctypes.CDLL(OUT)
P, I = ctypes.c_void_p, ctypes.c_int
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
N_true, D = int(os.environ.get("N", 640)), int(os.environ.get("D", 128))
N_cap = max(3600, 2 * N_true)
g = torch.Generator().manual_seed(0)
X = torch.randn(N_cap, D, generator=g).to(dev)
DQ = torch.randn(N_cap, 4 * D, generator=g).to(dev)
G = torch.randn(N_cap, D, generator=g).to(dev)
W = [torch.randn(D, D, generator=g).to(dev) * 0.1 for _ in range(4)]
b = [torch.randn(D, generator=g).to(dev) for _ in range(4)]
hi, lo = torch.empty(4 * D, D, device=dev), torch.empty(4 * D, D, device=dev)
hiT, loT = torch.empty(D, 4 * D, device=dev), torch.empty(D, 4 * D, device=dev)
A = torch.empty(D, D, device=dev)
nd = torch.tensor([N_true], dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
lib.ctan_wcat_prep.argtypes = [P] * 4 + [ctypes.c_float, I] + [P] * 5 + [P]
assert lib.ctan_wcat_prep(W[0].data_ptr(), W[1].data_ptr(), W[2].data_ptr(), W[3].data_ptr(), 0.1, D, hi.data_ptr(), lo.data_ptr(),
                          A.data_ptr(), hiT.data_ptr(), loT.data_ptr(), st) == 0
QK = torch.empty(N_cap, 4 * D, device=dev)
DX = torch.zeros(N_cap, D, device=dev)
lib.ctan_qkva_fwd_tc.argtypes = [P, I, P, I, P, P, P, P, P, P, P, I, P]
lib.ctan_dx_tc.argtypes = [P, P, I, P, I, P, P, P, I, I, P]
lib.ctan_tc_trace_read.argtypes = [P, I]
hint = (int(os.environ.get("HINT", 900)) << 8)


def qkva():
    return lib.ctan_qkva_fwd_tc(X.data_ptr(), N_cap, nd.data_ptr(), D, hi.data_ptr(), lo.data_ptr(), b[0].data_ptr(), b[1].data_ptr(),
                                b[2].data_ptr(), b[3].data_ptr(), QK.data_ptr(), hint, st)


def dx(split):
    return lambda: lib.ctan_dx_tc(DQ.data_ptr(), G.data_ptr(), N_cap, nd.data_ptr(), D, hiT.data_ptr(), loT.data_ptr(), DX.data_ptr(),
                                  hint, split, st)


NAMES = {0: "entry", 1: "setup", 2: "tma0", 4: "mma0", 5: "commit", 6: "acc", 9: "c0_ld", 10: "c0_stg", 11: "c0_st", 12: "c1_ld",
         13: "c1_stg", 14: "c1_st", 7: "epi", 8: "end"}
flush = torch.zeros(64 * 1024 * 1024, device=dev)
for name, fn in (("qkvA fwd [640,128]x[128,512]", qkva), ("dX split 1", dx(1)), ("dX split 2", dx(2)), ("dX split 4", dx(4))):
    for _ in range(5):
        assert fn() == 0
    torch.cuda.synchronize()
    for cold in (False, True):
        buf = (ctypes.c_ulonglong * (64 * 32))()
        lib.ctan_tc_trace_read(buf, 1)
        if cold:
            flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        lib.ctan_tc_trace_read(buf, 0)
        t = torch.tensor(list(buf), dtype=torch.int64).view(64, 32)
        work = [c for c in range(64) if int(t[c, 0]) > 0]
        t0 = min(int(t[c, 0]) for c in work)
        print(f"== {name}  {'L2 flushed' if cold else 'warm'}: {e0.elapsed_time(e1) * 1e3:.1f} us (events), {len(work)} working CTAs")
        for c in work[:3] + work[-1:]:
            row = {NAMES[k]: round((int(t[c, k]) - t0) / 1e3, 2) for k in NAMES if int(t[c, k]) > 0}
            kbs = [round((int(t[c, 16 + i]) - t0) / 1e3, 2) for i in range(12) if int(t[c, 16 + i]) > 0]
            print(f"   cta {c}: {row}  k-blocks landed at {kbs}")
