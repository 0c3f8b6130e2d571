"""Diagnostic: one training step from the SAME state (bench workload after STEPS optimizer steps, lr = 0 for the probed
step), three times: two-step forward, two-step forward again (the float-atomic noise floor of the backward), folded
forward.  Prints per parameter max|g_a - g_b| / max|g|."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CTAN_TRAIN_FUSE_ENC"] = "0"
import torch
import bench

dev = torch.device("cuda:0")
steps = int(os.environ.get("STEPS", "40"))
st, mean, std = bench.make_workload((steps + 8) * 200)
eng = bench.build_engine(st, 1, mean, std, dev)
m, ld = eng.model, eng.loader
eng.run(steps, train=True)
eng.set_hyper(lr=0.0)
torch.cuda.synchronize()


def snap():
    return [t.clone() for t in (m.memory.memory, m.memory.last_update, ld.e_id, ld.neighbors, ld._deg, eng.w["counters"])]


def restore(s):
    for t, v in zip((m.memory.memory, m.memory.last_update, ld.e_id, ld.neighbors, ld._deg, eng.w["counters"]), s):
        t.copy_(v)


def one(mode):
    eng.fuse_train = mode
    eng._graphs.clear()
    eng.run(1, train=True)
    torch.cuda.synchronize()
    return eng.flat_g.clone(), float(eng.losses(1)[0])


s0 = snap()
ga, la = one(False)
restore(s0)
gb, lb = one(False)
restore(s0)
gc, lc = one(True)
print("loss two-step", la, lb, "folded", lc)
off = 0
for name, p in zip([n for n, _ in m.named_parameters() if "lin_skip" not in n], eng._plist):
    pass
names = {id(p): n for n, p in m.named_parameters()}
for p in eng._plist:
    n = p.numel()
    a, b, c = ga[off:off + n], gb[off:off + n], gc[off:off + n]
    s = float(a.abs().max()) + 1e-30
    print(f"{names.get(id(p), '?'):36s} max|g| {s:.2e}   rerun {float((a - b).abs().max()) / s:.2e}   folded {float((a - c).abs().max()) / s:.2e}")
    off += n
