"""A few eager (non-graph) TGB-eval batches between cudaProfilerStart/Stop, for ncu.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_eval.csv python scripts/profile_eval.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench_tgb  # noqa: E402

dev = torch.device("cuda:0")
eng, st = bench_tgb.build_eval(0, 1, dev, 40)
eng.run(20)                       # graph replays: warm state + caches
eng.use_graph = False
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.run(3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
eng.check()
print("profiled 3 batches; N,E =", eng.w["counts"][:2].tolist(), "mrr", eng.mrr())
