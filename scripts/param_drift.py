"""Diagnostic: parameter drift of the engine against the oracle after n Adam steps on the bench workload (the quantity
tests/test_engine_gpu.py::test_bench_configuration_matches_oracle bounds), per parameter, plus the share of elements
whose oracle gradient history is at rounding level (Adam turns those into +-lr steps of arbitrary sign)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oracle import ctan_oracle as O

K = int(os.environ.get("ITERS", "1"))
steps = int(os.environ.get("STEPS", "42"))
dev = torch.device("cuda:0")
B, S = bench.WIKI["B"], bench.WIKI["S"]
st, mean, std = bench.make_workload((steps + 8) * B)
eng = bench.build_engine(st, K, mean, std, dev)
mm = eng.model
torch.manual_seed(0)
om = O.OracleCTAN(**bench.model_kwargs(st, K, mean, std))
lo_ref = O.TorchNeighborLoader(st["num_nodes"], S)
h = torch.zeros(st["num_nodes"], dtype=torch.long)
opt = torch.optim.Adam(om.parameters(), lr=bench.WIKI["lr"])
ref = [float(O.train_step_ref(om, lo_ref, h, opt, st, it * B, (it + 1) * B, st["neg"][it * B:(it + 1) * B])[0]) for it in range(steps)]
eng.run(steps, train=True)
torch.cuda.synchronize()
mine = eng.losses(steps).cpu().tolist()
print("max rel loss diff (last 12):", max(abs(a - b) / abs(b) for a, b in zip(mine[-12:], ref[-12:])))
po = dict(om.named_parameters())
for k, p in mm.named_parameters():
    if "lin_skip" in k:
        continue
    dlt = (p.detach().cpu() - po[k].detach()).abs()
    v = opt.state[po[k]]["exp_avg_sq"].sqrt()
    i = int(dlt.argmax())
    print(f"{k:40s} drift/max {float(dlt.max()) / float(po[k].abs().max()):.2e}   sqrt(v) at worst element "
          f"{float(v.flatten()[i]):.2e} (median {float(v.median()):.2e})")
