"""Full-size TGB evaluation: per-batch MRR of the GPU engine vs the reference's per-query loop (oracle port) on the
CPU, same stream / model / negatives / 1M-event warm state, first 3 batches after the warm insert."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # (test infrastructure: imports the oracle)
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench_tgb
from oracle import ctan_oracle as O
from ctan_b200 import synth

NB = 3
dev = torch.device("cuda:0")
eng, st = bench_tgb.build_eval(0, 1, dev, NB + 2)
gpu = []
for it in range(NB):
    eng.w["acc"].zero_()
    eng.run(1)
    gpu.append(eng.mrr())
cfg, B, warm = bench_tgb.COMMENT, bench_tgb.COMMENT["B"], bench_tgb.WARM_EVENTS
g = torch.Generator().manual_seed(1)
neg = torch.randint(st["min_dst"], st["max_dst"] + 1, (st["src"].numel(), cfg["n_neg"]), generator=g)
assert torch.equal(neg, eng.neg.cpu())
mean, std = synth.delta_t_stats(st, 20000)
torch.manual_seed(0)
om = O.OracleCTAN(num_nodes=st["num_nodes"], edge_dim=st["msg"].size(1), memory_dim=cfg["D"], time_dim=cfg["T"], node_dim=1,
                  num_iters=cfg["K"], epsilon=cfg["eps"], gamma=cfg["gamma"], mean_delta_t=mean, std_delta_t=std, tgb=True)
loader = O.TorchNeighborLoader(st["num_nodes"], cfg["S"])
for lo in range(0, warm, 2000):
    loader.insert(st["src"][lo:lo + 2000], st["dst"][lo:lo + 2000])
gen = torch.Generator(device="cpu").manual_seed(7)
touched = torch.unique(torch.cat([st["src"][:warm], st["dst"][:warm]]))
om.memory.memory[touched] = 0.1 * torch.randn(touched.numel(), cfg["D"], generator=gen)
om.memory.last_update[touched] = st["t"][warm - 1] // 2
helper = torch.zeros(st["num_nodes"], dtype=torch.long)
cpu = []
with torch.no_grad():
    for it in range(NB):
        lo, hi = warm + it * B, warm + (it + 1) * B
        m, _, _ = O.tgb_eval_batch_ref(om, loader, helper, st, lo, hi, [neg[q].tolist() for q in range(lo, hi)])
        cpu.append(float(np.mean(m)))
print("gpu per-batch MRR", gpu)
print("cpu per-batch MRR", cpu)
print("max |diff|", max(abs(a - b) for a, b in zip(gpu, cpu)))
assert max(abs(a - b) for a, b in zip(gpu, cpu)) < 5e-3
assert torch.equal(eng.loader.e_id.cpu(), loader.e_id)
mo = om.memory.memory
rel = float((eng.model.memory.memory.cpu() - mo).abs().max() / mo.abs().max())
print("memory max rel diff", rel)
assert rel < 1e-4
assert torch.equal(eng.model.memory.last_update.cpu(), om.memory.last_update)
