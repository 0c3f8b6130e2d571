"""A/B of whole-step time (graph replay, back-to-back) under environment switches."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, bench
    dev = torch.device("cuda:0")
    st, mean, std = bench.make_workload(900 * 200)
    eng = bench.build_engine(st, int(os.environ.get("ITERS", "1")), mean, std, dev)
    eng.run(300, train=True); torch.cuda.synchronize()
    ms = bench.time_back_to_back(eng, 300, True)
    eng.reset(); eng.run(100, train=False); torch.cuda.synchronize()
    msi = bench.time_back_to_back(eng, 300, False)
    print(f"train {ms/300*1e3:7.1f} us/step   infer {msi/300*1e3:7.1f} us/step")
else:
    for env in [{}, {"CTAN_TC_DX": "0"}, {"CTAN_NO_TC": "1"}, {"CTAN_TC_MODE": "1"}]:
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
        print(env, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
