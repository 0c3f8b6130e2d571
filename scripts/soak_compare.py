"""Soak: N training steps of the bench workload with the default engine (all round-2 restructurings on) against the
same engine with every A/B switch off (the reference's operation order, no early / deferred branches), and against a
second default run (the float-atomic noise floor).  A race or a stale-buffer bug shows up as a jump in the loss gap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

SW = ("CTAN_TRAIN_FUSE_ENC", "CTAN_EPROJ_ROWS", "CTAN_DEFER_PREP", "CTAN_DX_FIRST", "CTAN_WGRAD_DIRECT", "CTAN_TENC_FUSED",
      "CTAN_LINK_FUSED", "CTAN_TRAIN_EARLY", "CTAN_DQ_MULTI", "CTAN_INFER_FUSE_ENC")
dev = torch.device("cuda:0")
n = int(os.environ.get("STEPS", "600"))
K = int(os.environ.get("ITERS", "1"))
st, mean, std = bench.make_workload((n + 8) * 200)


def run(off):
    for k in SW:
        os.environ[k] = "0" if off else "1"
    eng = bench.build_engine(st, K, mean, std, dev)
    for chunk in (n // 3, n // 3, n - 2 * (n // 3)):           # several run() calls: the pipelined sequence restarts
        eng.run(chunk, train=True)
    torch.cuda.synchronize()
    eng.check()
    return eng.losses(n).cpu().double(), eng.model.memory.memory.clone()


a, ma = run(False)
b, mb = run(False)
c, mc = run(True)
assert torch.isfinite(a).all() and torch.isfinite(c).all()
for name, x, y in (("default vs default", a, b), ("default vs all switches off", a, c)):
    d = (x - y).abs() / y.abs()
    print(f"{name}: rel loss gap  first 50 max {float(d[:50].max()):.2e}   all max {float(d.max()):.2e}   mean {float(d.mean()):.2e}"
          f"   last 50 mean {float(d[-50:].mean()):.2e}")
print("memory gap default vs off:", float((ma - mc).abs().max()) / float(mc.abs().max()),
      " default vs default:", float((ma - mb).abs().max()) / float(mb.abs().max()))
print("mean loss last 100: default", float(a[-100:].mean()), " switches off", float(c[-100:].mean()))
