"""Host-side profile of the drop-in path: the reference's loop body on ctan_b200 modules (cProfile, 200 steps)."""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oracle import ctan_oracle as O
from ctan_b200.models import CTAN, LastNeighborLoader
dev = "cuda:0"
cfg = bench.WIKI
st, mean, std = bench.make_workload(300 * 200)
torch.manual_seed(0)
om = CTAN(**bench.model_kwargs(st, 1, mean, std)).to(dev)
ld = LastNeighborLoader(st["num_nodes"], cfg["S"], device=dev)
h = torch.zeros(st["num_nodes"], dtype=torch.long, device=dev)
opt = torch.optim.Adam(om.parameters(), lr=cfg["lr"])
g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in st.items()}
B = cfg["B"]
for it in range(30):
    O.train_step_ref(om, ld, h, opt, g, it * B, (it + 1) * B, g["neg"][it * B:(it + 1) * B])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for it in range(30, 230):
    O.train_step_ref(om, ld, h, opt, g, it * B, (it + 1) * B, g["neg"][it * B:(it + 1) * B])
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
