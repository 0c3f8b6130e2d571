"""Host-side profile (cProfile) of the reference's own train loop running UNMODIFIED on ctan_b200.models: where the
1.8 ms per step of the module path go."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

st, mean, std = bench.make_workload(520 * 200)
bench.reference_train_rate(st, 1, mean, std, "cuda:0", 20, 10, ours=True)            # load kernels / caches
pr = cProfile.Profile()
pr.enable()
r = bench.reference_train_rate(st, 1, mean, std, "cuda:0", 200, 30, ours=True)
pr.disable()
print(r)
ps = pstats.Stats(pr)
ps.sort_stats("tottime").print_stats(35)
