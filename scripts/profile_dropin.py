"""Host-side profile of the drop-in path: the reference's loop body on ctan_b200 modules (cProfile, 200 steps)."""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
st, mean, std = bench.make_workload(300 * 200)
bench.reference_train_rate(st, 1, mean, std, "cuda:0", 20, 10, ours=True)            # load kernels / caches
pr = cProfile.Profile()
pr.enable()
print(bench.reference_train_rate(st, 1, mean, std, "cuda:0", 200, 30, ours=True))
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
