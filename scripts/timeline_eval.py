"""Device timeline of one TGB-evaluation batch inside the CUDA graph (tgbl-comment-shaped workload of bench_tgb.py, one
rank): a %globaltimer probe after every call, on every stream (issue order).  Front-end done beforehand, as run(n) has it
(prefetched beside the previous batch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench_tgb
from ctan_b200 import _lib

dev = torch.device("cuda:0")
eng, st = bench_tgb.build_eval(0, 1, dev, 60)
eng.run(30)
torch.cuda.synchronize()
W = int(os.environ.get("AS_WORLD", "1"))
if W > 1:     # the shard rank 0 of a W-rank run scores (single process: the exchange itself is not in this timeline)
    from ctan_b200.tgb_eval import shard_range
    eng.q0, eng.q1 = shard_range(eng.B, 0, W)
    eng.Q = eng.q1 - eng.q0
    seeds = eng.Q * (2 + eng.n_neg)
    eng._n_hint = min(eng.n_cap, 2 * seeds) << 8
    eng._e_hint = min(eng.e_cap, 4 * seeds) << 8
with torch.cuda.device(dev):
    eng._frontend(0)
torch.cuda.synchronize()
ts = torch.zeros(256, dtype=torch.int64, device=dev)
names = []
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    _lib.lib().ctan_probe(ts.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    names.append("start")
    _lib.PROBE = (ts, names, torch.cuda.current_stream().cuda_stream)
    _lib.PROBE_SIDE = True
    try:
        eng._batch(0, os.environ.get("PREFETCH", "0") == "1")
        eng._update(0, True, 0)
    finally:
        _lib.PROBE = None
        _lib.PROBE_SIDE = False
R = 50
acc = torch.zeros(len(names), dtype=torch.float64)
for _ in range(R):
    g.replay()
    torch.cuda.synchronize()
    t = ts[:len(names)].cpu().double()
    acc += (t - t[0])
acc /= R
for n, a in zip(names, acc.tolist()):
    print(f"{a/1e3:9.2f} us  after {n}")
