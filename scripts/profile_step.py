"""Run a few eager (non-graph) engine steps between cudaProfilerStart/Stop, for ncu.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 3
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warm", type=int, default=300)
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--infer", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
st, mean, std = bench.make_workload((a.warm + a.steps + 8) * 200)
eng = bench.build_engine(st, a.iters, mean, std, dev)
eng.use_graph = False
eng.run(a.warm, train=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.run(a.steps, train=not a.infer)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
eng.check()
print("profiled", a.steps, "steps; N,E =", eng.w["counts"][:2].tolist())
