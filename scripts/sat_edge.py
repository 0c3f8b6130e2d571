"""Saturating-shape roofline of the fused edge kernel (attention logits + segmented softmax + aggregation +
tanh update): N destination rows of degree S over a node table much larger than L2.  Achieved HBM GB/s =
ALGORITHMIC bytes / CUDA-event time.  Algorithmic bytes per launch:
   per node : q, xA+b, x_in read + x_out written            = 16*D bytes
   per edge : k_j, v_j gathers + EP row (read once) + src id  = 12*D + 8 bytes"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ctan_b200
from ctan_b200._lib import call, ptr


def run(N=100_000, S=16, D=256, iters=10, dev=torch.device("cuda:0")):
    E = N * S
    g = torch.Generator(device="cpu").manual_seed(0)
    QKVA = torch.randn(N, 4 * D, device=dev)
    EP = torch.randn(E, D, device=dev)
    X = torch.randn(N, D, device=dev)
    Xo = torch.empty(N, D, device=dev)
    rowptr = (torch.arange(N + 1, dtype=torch.int32) * S).to(dev)
    esrc = torch.randint(0, N, (E,), generator=g).to(dev)
    fn = lambda: call("ctan_edge_fwd", ptr(QKVA), ptr(EP), ptr(rowptr), ptr(esrc), N, None, D, ptr(X), 0.5, ptr(Xo),
                      None, None, None, None, 0, 0)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    alg = N * 16 * D + E * (12 * D + 8)
    return {"kernel": "edge_fwd_kernel", "N": N, "E": E, "D": D, "ms_per_launch": ms, "algorithmic_bytes": alg,
            "achieved_gbs": alg / (ms * 1e-3) / 1e9}


if __name__ == "__main__":
    r = run(N=int(os.environ.get("SAT_N", 100000)))
    peak = 6545.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    r["peak_gbs"] = peak
    r["frac"] = r["achieved_gbs"] / peak
    print(json.dumps(r))
