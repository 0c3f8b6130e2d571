"""Diagnostic: accuracy of the folded enc_x + projection contraction on REAL training state (bench workload, after
STEPS optimizer steps): X0 / QKVA of the two-step path (what the step just computed, lr set to 0 for that step so the
weights are the ones it used) and of the folded path, both against fp64 from the same gathered rows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CTAN_TRAIN_FUSE_ENC"] = "0"
import torch
import bench
from ctan_b200._lib import call, ptr

dev = torch.device("cuda:0")
steps = int(os.environ.get("STEPS", "40"))
st, mean, std = bench.make_workload((steps + 8) * 200)
eng = bench.build_engine(st, 1, mean, std, dev)
eng.run(steps, train=True)
eng.set_hyper(lr=0.0)
eng.run(1, train=True)
torch.cuda.synchronize()
w, m, D, Fn = eng.w, eng.model, eng.D, eng.Fn
N = int(eng.fe[(steps) & 1]["Nd"].item()) if eng.prefetch else int(eng.fe[0]["Nd"].item())
z0 = w["z0"][:N].clone()                                   # [N, D+Fn] gathered rows of that step
X0_a, Q_a = w["X"][0][:N].clone(), w["QKVA"][0][:N].clone()
g = m.gnn
ph = g.aconv.phi
Wenc, benc = g.enc_x.weight.detach(), g.enc_x.bias.detach()
Ws = [ph.lin_query.weight.detach(), ph.lin_key.weight.detach(), ph.lin_value.weight.detach(), g.aconv.W.detach()]
bs = [ph.lin_query.bias.detach(), ph.lin_key.bias.detach(), ph.lin_value.bias.detach(), g.aconv.bias.detach()]
gamma = float(g.aconv.gamma)
A = Ws[3] - Ws[3].t() - gamma * torch.eye(D, device=dev)
Wcat = torch.cat([Ws[0], Ws[1], Ws[2], A]).double()
X0_ref = z0.double() @ Wenc.double().t() + benc.double()
Q_ref = X0_ref @ Wcat.t() + torch.cat(bs).double()
Kx = ((D + Fn + 31) // 32) * 32
hi, lo, bce = torch.empty(5 * D, Kx, device=dev), torch.empty(5 * D, Kx, device=dev), torch.empty(5 * D, device=dev)
call("ctan_fuse_enc_prep_split", ptr(Ws[0]), ptr(Ws[1]), ptr(Ws[2]), ptr(Ws[3]), gamma, ptr(bs[0]), ptr(bs[1]), ptr(bs[2]),
     ptr(bs[3]), ptr(Wenc), ptr(benc), D, Fn, Kx, ptr(hi), ptr(lo), ptr(bce))
zext = torch.zeros(N + 128, Kx, device=dev)
zext[:N, :D + Fn] = z0
X0_b, Q_b = torch.empty(N, D, device=dev), torch.empty(N, 4 * D, device=dev)
call("ctan_enc_qkva_tc", ptr(zext), N, None, Kx, ptr(hi), ptr(lo), D, ptr(bce), ptr(X0_b), ptr(Q_b), 0)
torch.cuda.synchronize()
print("N", N, "|z| max", float(z0.abs().max()), "|X0| max", float(X0_ref.abs().max()), "|Q| max", float(Q_ref.abs().max()),
      "|Wenc| max", float(Wenc.abs().max()), "|Wcat| max", float(Wcat.abs().max()))
for name, a, b, r in (("X0", X0_a, X0_b, X0_ref), ("QKVA", Q_a, Q_b, Q_ref)):
    s = float(r.abs().max())
    print(f"{name}: two-step err/max {float((a.double() - r).abs().max()) / s:.2e}   folded err/max {float((b.double() - r).abs().max()) / s:.2e}"
          f"   two-step vs folded {float((a - b).abs().max()) / s:.2e}")
for i, nm in enumerate(("q", "k", "v", "xA")):
    r = Q_ref[:, i * D:(i + 1) * D]
    s = float(r.abs().max())
    print(f"  {nm}: |ref| max {s:.3e}  two-step {float((Q_a[:, i*D:(i+1)*D].double() - r).abs().max()) / s:.2e}  folded {float((Q_b[:, i*D:(i+1)*D].double() - r).abs().max()) / s:.2e}")
