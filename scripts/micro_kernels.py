"""Per-kernel warm timings at the wiki-config sizes: each C-ABI call replayed 50x inside a CUDA graph
(L2-warm, no host overhead).  Used to tune the kernels; not a benchmark."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ctan_b200 import _lib
from ctan_b200._lib import call, ptr

dev = torch.device("cuda:0")
st, mean, std = bench.make_workload(400 * 200)
eng = bench.build_engine(st, int(os.environ.get("ITERS", "1")), mean, std, dev)
eng.run(300, train=True)
torch.cuda.synchronize()
w, m, ld = eng.w, eng.model, eng.loader
print("N,E =", w["counts"][:2].tolist(), "tile", os.environ.get("CTAN_GEMM_TILE"))
N, E, D, Fe, Fn, T, Hd, O, P, B, S, K = eng.n_cap, eng.e_cap, eng.D, eng.Fe, eng.Fn, eng.T, eng.Hd, eng.O, eng.P, eng.B, eng.S, eng.K
Nd, Ed = ptr(eng._Nd), ptr(eng._Ed)
(Wenc, benc, Wq, bq, Wk, bk, Wv, bv, We, W, b, tw, tb, Wsrc, bsrc, Wdst, bdst, Wfin, bfin) = (ptr(t) for t in eng.p)
(gWenc, gbenc, gWq, gbq, gWk, gbk, gWv, gbv, gWe, gW, gb, gtw, gtb, gWsrc, gbsrc, gWdst, gbdst, gWfin, gbfin) = (ptr(t) for t in eng.g)
mem, lu = m.memory.memory, m.memory.last_update
z = w["X"][K]
xf = ptr(eng.x)
tests = {
    "trivial(antisym_prep D=1)": lambda: call("ctan_antisym_prep", W, 0.1, 1, ptr(w["A"])),
    "antisym_prep": lambda: call("ctan_antisym_prep", W, 0.1, D, ptr(w["A"])),
    "enc_fwd(gather)": lambda: call("ctan_enc_fwd", ptr(mem), D, xf, Fn, ptr(w["n_id"]), None, N, Nd, Wenc, benc, ptr(w["X"][0])),
    "eproj_fwd": lambda: call("ctan_eproj_fwd", ptr(eng.msg), Fe, ptr(w["eid"]), ptr(w["rel"]), tw, tb, T, We, E, Ed, D, ptr(w["EP"])),
    "qkva_fwd": lambda: call("ctan_qkva_fwd", ptr(w["X"][0]), N, Nd, D, Wq, Wk, Wv, ptr(w["A"]), bq, bk, bv, b, ptr(w["QKVA"][0])),
    "edge_fwd": lambda: call("ctan_edge_fwd", ptr(w["QKVA"][0]), ptr(w["EP"]), ptr(w["rowptr"]), ptr(w["esrc"]), N, Nd, D, ptr(w["X"][0]), 0.1, ptr(w["X"][1]), ptr(w["TH"][0]), ptr(w["ALPHA"][0]), None),
    "readout_fwd(2k)": lambda: call("ctan_readout_fwd", ptr(z), D, ptr(w["pair_src"]), ptr(w["pair_dst"]), ptr(ld._assoc), P, None, Hd, O, Wsrc, bsrc, Wdst, bdst, Wfin, bfin, ptr(w["H"]), ptr(w["OUT"]), 1),
    "readout_bwd dz(3k)": lambda: call("ctan_readout_bwd", ptr(w["dOUT"]), ptr(w["H"]), ptr(z), D, ptr(w["pair_src"]), ptr(w["pair_dst"]), ptr(ld._assoc), P, None, Hd, O, Wsrc, Wdst, Wfin, ptr(w["dHid"]), ptr(w["dZ"]), None, None, None, None, None, None),
    "readout_bwd w(2k)": lambda: call("ctan_readout_bwd", ptr(w["dOUT"]), ptr(w["H"]), ptr(z), D, ptr(w["pair_src"]), ptr(w["pair_dst"]), ptr(ld._assoc), P, None, Hd, O, Wsrc, Wdst, Wfin, ptr(w["dHid"]), None, gWsrc, gbsrc, gWdst, gbdst, gWfin, gbfin),
    "edge_bwd": lambda: call("ctan_edge_bwd", ptr(w["QKVA"][0]), ptr(w["EP"]), ptr(w["rowptr"]), ptr(w["esrc"]), N, Nd, D, ptr(w["dZ"]), ptr(w["TH"][0]), ptr(w["ALPHA"][0]), 0.1, ptr(w["DQKVA"]), ptr(w["dEP"]), 0, ptr(w["DS"])),
    "qkva_bwd dx": lambda: call("ctan_qkva_bwd", ptr(w["DQKVA"]), ptr(w["X"][0]), ptr(w["dZ"]), N, Nd, D, Wq, Wk, Wv, ptr(w["A"]), ptr(eng._DX[0]), None, None, None, None, None, None, None, None, eng.n_typ),
    "qkva_bwd dW": lambda: call("ctan_qkva_bwd", ptr(w["DQKVA"]), ptr(w["X"][0]), ptr(w["dZ"]), N, Nd, D, Wq, Wk, Wv, ptr(w["A"]), None, gWq, gWk, gWv, ptr(w["dA"]), gbq, gbk, gbv, gb, eng.n_typ),
    "enc_bwd": lambda: call("ctan_enc_bwd", ptr(eng._DX[0]), ptr(w["z0"]), N, Nd, D, Fn, gWenc, gbenc, eng.n_typ),
    "eproj_bwd(2k)": lambda: call("ctan_eproj_bwd", ptr(w["dEP"]), ptr(eng.msg), Fe, ptr(w["eid"]), ptr(w["rel"]), tw, tb, T, We, E, Ed, D, gWe, ptr(w["dTenc"]), eng.e_typ),
    "tenc_bwd": lambda: call("ctan_tenc_bwd", ptr(w["dTenc"]), ptr(w["rel"]), tw, tb, E, Ed, T, gtw, gtb),
    "adam": lambda: call("ctan_adam_step", ptr(eng.flat_p), ptr(eng.flat_g), ptr(eng.adam_m), ptr(eng.adam_v), eng.n_params, 0.0, 0.9, 0.999, 1e-8, 0.0, 1, None),
    "memset zero_buf": lambda: eng.zero_buf.zero_(),
}
R = 50
for name, fn in tests.items():
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(R):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / (4 * R) * 1e3:8.2f} us/call")
