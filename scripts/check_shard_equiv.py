"""Run under torchrun (N ranks): the sharded TGB evaluation must reproduce the 1-rank run exactly
(bitwise memory / neighbour state, identical MRR).  Each rank also runs a private world=1 engine on the
same stream as its own reference."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench_tgb

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(bench_tgb.COMMENT, D=64, T=32, S=8)
nb = 12
eng, st = bench_tgb.build_eval(rank, world, dev, nb, cfg=cfg, n_nodes_scale=0.01, warm=20000)
eng.run(nb)
torch.cuda.synchronize()
mem_s, lu_s, eid_s = eng.model.memory.memory.clone(), eng.model.memory.last_update.clone(), eng.loader.e_id.clone()
mrr_s = eng.mrr()
ref, _ = bench_tgb.build_eval(0, 1, dev, nb, cfg=cfg, n_nodes_scale=0.01, warm=20000)
ref.run(nb)
torch.cuda.synchronize()
ok = (torch.equal(mem_s, ref.model.memory.memory) and torch.equal(lu_s, ref.model.memory.last_update)
      and torch.equal(eid_s, ref.loader.e_id) and abs(mrr_s - ref.mrr()) < 1e-9)
flags = [None] * world
dist.all_gather_object(flags, bool(ok))
if rank == 0:
    print("shard-equivalence", "OK" if all(flags) else "MISMATCH", flags, "mrr", mrr_s, ref.mrr(), "graph", eng._graph is not None, "mem", torch.equal(mem_s, ref.model.memory.memory), "lu", torch.equal(lu_s, ref.model.memory.last_update), "eid", torch.equal(eid_s, ref.loader.e_id), flush=True)
eng.close()
dist.destroy_process_group()
sys.exit(0 if all(flags) else 1)
