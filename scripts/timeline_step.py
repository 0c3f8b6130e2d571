"""Device timeline of one training step inside the CUDA graph: a probe kernel (globaltimer) after every
main-stream call.  Prints the mean gap between consecutive probes over R replays."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ctan_b200 import _lib

dev = torch.device("cuda:0")
st, mean, std = bench.make_workload(800 * 200)
eng = bench.build_engine(st, int(os.environ.get("ITERS", "1")), mean, std, dev)
eng.run(300, train=True)
torch.cuda.synchronize()
if os.environ.get("PREFETCHED", "0") == "1":
    with torch.cuda.device(dev):
        eng._frontend(0, with_rel=os.environ.get("EARLY", "0") != "1")
    torch.cuda.synchronize()
train = os.environ.get("INFER", "0") != "1"
ts = torch.zeros(256, dtype=torch.int64, device=dev)
names = []
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    _lib.lib().ctan_probe(ts.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    names.append("start")
    _lib.PROBE = (ts, names, torch.cuda.current_stream().cuda_stream)
    _lib.PROBE_SIDE = os.environ.get("ALL", "0") == "1"      # also probe the side-stream calls (order = issue order)
    if os.environ.get("PREFETCHED", "0") == "1":             # the step as run(n) replays it: front-end done beforehand
        try:
            eng._step(train, 0, os.environ.get("PF", "0") == "1", early=os.environ.get("EARLY", "0") == "1")
        finally:
            _lib.lib().ctan_set_pdl(1, None)
    else:
        eng._launch(train)
    _lib.PROBE = None
R = 100
acc = torch.zeros(len(names), dtype=torch.float64)
tot = 0.0
for _ in range(R):
    g.replay()
    torch.cuda.synchronize()
    t = ts[:len(names)].cpu().double()
    acc += (t - t[0])
    tot += float(t[-1] - t[0])
acc /= R
prev = 0.0
for n, a in zip(names, acc.tolist()):
    print(f"{a/1e3:9.2f} us  (+{(a-prev)/1e3:7.2f})  after {n}")
    prev = a
print("probes", len(names), "total", tot / R / 1e3, "us (includes ~1.5 us per probe)")
