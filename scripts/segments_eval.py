"""Per-batch GPU time of the three segments of a sharded TGB-eval batch (graph A: score + prefetch | all-gather |
graph B: memory write-back), CUDA events between the launches, mean over 200 batches.  torchrun or single process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench_tgb
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NB = 200
eng, st = bench_tgb.build_eval(rank, world, dev, NB + 30)
eng.run(20)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(NB)]
if eng.prefetch:
    eng._graph.replay()
par = 0
for j in range(NB):
    ga, gb = eng._graphs_for(par, eng.prefetch and j < NB - 1)
    ev[j][0].record(); ga.replay(); ev[j][1].record()
    if world > 1:
        eng._exchange(); ev[j][2].record(); gb.replay()
    else:
        ev[j][2].record()
    ev[j][3].record()
    if eng.prefetch:
        par ^= 1
torch.cuda.synchronize()
eng.loader.cur_e_id += NB * eng.B
a = sum(e[0].elapsed_time(e[1]) for e in ev[10:]) / (NB - 10) * 1e3
x = sum(e[1].elapsed_time(e[2]) for e in ev[10:]) / (NB - 10) * 1e3
b = sum(e[2].elapsed_time(e[3]) for e in ev[10:]) / (NB - 10) * 1e3
tot = ev[10][0].elapsed_time(ev[-1][3]) / (NB - 11) * 1e3
print(f"rank {rank}/{world} prefetch={eng.prefetch}: graph A {a:.1f} us | exchange {x:.1f} us | graph B {b:.1f} us | per batch {tot:.1f} us", flush=True)
if world > 1:
    dist.destroy_process_group()
