"""Experiment: per-batch time of the scoring shard rank 0 of a W-rank run would process (single process, no exchange),
with the next batch's front-end inline vs prefetched beside the scoring."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench_tgb
from ctan_b200.tgb_eval import shard_range

dev = torch.device("cuda:0")
n = 300
for W in (1, 2, 8):
    for pf in (False, True):
        eng, st = bench_tgb.build_eval(0, 1, dev, 2 * n + 20)
        eng.q0, eng.q1 = shard_range(eng.B, 0, W)
        eng.Q = eng.q1 - eng.q0
        seeds = eng.Q * (2 + eng.n_neg)
        eng._n_hint = min(eng.n_cap, 2 * seeds) << 8
        eng._e_hint = min(eng.e_cap, 4 * seeds) << 8
        eng.prefetch = pf
        eng._graphs, eng._graph = {}, None
        eng.run(20)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run(n)
        e1.record()
        torch.cuda.synchronize()
        print(f"as world {W} prefetch {int(pf)}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per batch", flush=True)
        del eng
