"""Device timeline of one TGB-eval batch (1 GPU) inside the CUDA graph (globaltimer probes)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench_tgb
from ctan_b200 import _lib
dev = torch.device("cuda:0")
eng, st = bench_tgb.build_eval(0, 1, dev, 200)
eng.run(20); torch.cuda.synchronize()
ts = torch.zeros(256, dtype=torch.int64, device=dev); names = []
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    _lib.lib().ctan_probe(ts.data_ptr(), 0, torch.cuda.current_stream().cuda_stream); names.append("start")
    _lib.PROBE = (ts, names, torch.cuda.current_stream().cuda_stream)
    eng._frontend(0); eng._batch(0, False); eng._update(0)
    _lib.PROBE = None
acc = torch.zeros(len(names), dtype=torch.float64)
R = 50
for _ in range(R):
    g.replay(); torch.cuda.synchronize()
    t = ts[:len(names)].cpu().double(); acc += t - t[0]
acc /= R; prev = 0
for n, a in zip(names, acc.tolist()):
    print(f"{a/1e3:9.2f} us (+{(a-prev)/1e3:7.2f}) after {n}"); prev = a
print("N,E", eng.w["counts"][:2].tolist())
