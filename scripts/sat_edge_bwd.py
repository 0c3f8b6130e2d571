"""Saturating-shape roofline of the fused edge BACKWARD (scatter) kernel: N destination rows of degree S.
Algorithmic bytes per launch:
   per node : TH, G read; q read; dh (DQKVA[:,3D:]) and dq (DQKVA[:,:D]) written            = 20*D bytes
   per edge : k_j, v_j rows read (v_j twice: two passes over the row's edges -> counted once), EP row read,
              dk_j / dv_j atomically accumulated (read-modify-write: 2 x 8*D), dEP row written   = 12*D + 16*D + 4*D"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ctan_b200
from ctan_b200._lib import call, ptr


def run(N=100_000, S=16, D=256, iters=5, dev=torch.device("cuda:0")):
    E = N * S
    g = torch.Generator(device="cpu").manual_seed(0)
    QKVA = torch.randn(N, 4 * D, device=dev)
    EP = torch.randn(E, D, device=dev)
    G = torch.randn(N, D, device=dev)
    TH = torch.tanh(torch.randn(N, D, device=dev))
    ALPHA = torch.full((E,), 1.0 / S, device=dev)
    DQ = torch.zeros(N, 4 * D, device=dev)
    DEP = torch.empty(E, D, device=dev)
    DS = torch.empty(E, device=dev)
    rowptr = (torch.arange(N + 1, dtype=torch.int32) * S).to(dev)
    esrc = torch.randint(0, N, (E,), generator=g).to(dev)
    fn = lambda: call("ctan_edge_bwd", ptr(QKVA), ptr(EP), ptr(rowptr), ptr(esrc), N, None, D, ptr(G), ptr(TH), ptr(ALPHA),
                      0.5, ptr(DQ), ptr(DEP), 0, ptr(DS))
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    alg = N * 20 * D + E * (32 * D + 8)
    return {"kernel": "edge_bwd_kernel", "N": N, "E": E, "D": D, "ms_per_launch": ms, "algorithmic_bytes": alg,
            "achieved_gbs": alg / (ms * 1e-3) / 1e9}


if __name__ == "__main__":
    r = run()
    peak = 6545.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    r["peak_gbs"] = peak
    r["frac"] = r["achieved_gbs"] / peak
    print(json.dumps(r))
