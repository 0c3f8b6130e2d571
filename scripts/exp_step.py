"""Experiment: where does the step time go?  eager vs graph, single stream vs fork/join DAG."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ctan_b200 import _lib

dev = torch.device("cuda:0")
st, mean, std = bench.make_workload(1400 * 200)
for parallel in (False, True):
    for graph in (False, True):
        eng = bench.build_engine(st, 1, mean, std, dev)
        eng.parallel = parallel
        eng.use_graph = graph
        eng._graphs = {}
        eng.run(300, train=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ms = bench.time_back_to_back(eng, 300, True)
        wall = time.perf_counter() - t0
        print(f"parallel={parallel} graph={graph}: {ms/300*1e3:.1f} us/step device, wall {wall/300*1e6:.1f} us/step")
        if not parallel and not graph:
            prof = bench.per_kernel_profile(eng, 50, True)
            tot = 0
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms_per_call"] * kv[1]["calls_per_step"]):
                c = v["ms_per_call"] * v["calls_per_step"] * 1e3
                tot += c
                print(f"   {k:22s} {c:7.1f} us/step ({v['calls_per_step']:.0f} calls)")
            print("   sum", tot)
# empty-kernel graph chain: per-node overhead
x = torch.zeros(1, device=dev)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(40):
        _lib.call("ctan_antisym_prep", x.data_ptr(), 0.0, 1, x.data_ptr())
for _ in range(5): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): g.replay()
e1.record(); torch.cuda.synchronize()
print("40 trivial kernels in a graph:", e0.elapsed_time(e1) / 100 * 1e3, "us per replay")
