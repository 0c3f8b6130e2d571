"""TGB one-vs-many evaluation benchmark (BASELINE.json configs[4]: tgbl-comment-shaped stream, 994 790
nodes, 2-d messages, 20 negatives per query, CTAN_tgb D=T=256, S=32, K=1), sharded over N ranks.

Launched by bench.py (directly for N=1, under torchrun for N>1).  A "step" is one 200-event batch =
200 queries; value = queries/sec of the whole job (strong scaling: the batch is split over the ranks),
timed on the device between two barriers, max over ranks."""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch

COMMENT = dict(D=256, T=256, S=32, K=1, B=200, n_neg=20, eps=1.0, gamma=0.1, shape="comment", warm=1_000_000)
# tgbl-wiki variant (confs/conf_tgb_link_prediction.py:5-16: D=T=256, S=32, K in {1,2,3}; 172-d messages; the TGB
# validation set ranks each positive against ~999 negatives): the one reference config where a 200-query batch
# carries enough scoring work (200 k pairs) for query sharding to have something to split
WIKI_TGB = dict(D=256, T=256, S=32, K=1, B=200, n_neg=999, eps=1.0, gamma=0.1, shape="wiki", warm=50_000)
WARM_EVENTS = COMMENT["warm"]   # events inserted into the neighbour/memory state before the timed window

WORKLOAD = ("tgbl-comment-shaped TGB one-vs-many eval: 994790 nodes, 2-d msgs, 200 queries/batch x 20 negatives, "
            "CTAN_tgb D=T=256 S=32 K=1; timed window follows a 1000000-event warm insert (SURVEY.md S8d)")
L2_NOTE = "state tables (1.0 GB memory + 0.5 GB neighbour rows) exceed L2; no flush"


def build_eval(rank, world, dev, n_batches, shape=None, cfg=COMMENT, n_nodes_scale=1.0, warm=None):
    import ctan_b200  # noqa: F401
    from ctan_b200 import synth
    from ctan_b200._lib import call, ptr
    from ctan_b200.models import CTAN_tgb, LastNeighborLoader
    from ctan_b200.tgb_eval import TGBEvalEngine
    B = cfg["B"]
    shape = shape or cfg.get("shape", "comment")
    warm = cfg.get("warm", WARM_EVENTS) if warm is None else warm
    n_events = warm + (n_batches + 4) * B
    st = synth.make_stream(shape, n_events=n_events, seed=0, tgb=True, n_nodes_scale=n_nodes_scale)
    g = torch.Generator().manual_seed(1)
    neg = torch.randint(st["min_dst"], st["max_dst"] + 1, (n_events, cfg["n_neg"]), generator=g)
    mean, std = synth.delta_t_stats(st, 20000)
    torch.manual_seed(0)
    model = CTAN_tgb(num_nodes=st["num_nodes"], edge_dim=st["msg"].size(1), memory_dim=cfg["D"], time_dim=cfg["T"],
                     node_dim=1, num_iters=cfg["K"], epsilon=cfg["eps"], gamma=cfg["gamma"], mean_delta_t=mean,
                     std_delta_t=std).to(dev)
    loader = LastNeighborLoader(st["num_nodes"], cfg["S"], device=dev)
    gs = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in st.items()}
    gs["neg"] = neg.to(dev)
    group = None
    eng = TGBEvalEngine(model, loader, gs, batch_size=B, n_neg=cfg["n_neg"], rank=rank, world=world, group=group)
    eng.warmup()
    # warm state: insert the first `warm` events and give the touched nodes a non-trivial memory
    # (replicated identically on every rank; deterministic)
    model.reset_memory()
    loader.reset_state()
    chunk = 2000
    ws = torch.zeros(4 * chunk, dtype=torch.int32, device=dev)
    for lo in range(0, warm, chunk):
        loader.insert(gs["src"][lo:lo + chunk], gs["dst"][lo:lo + chunk])
    gen = torch.Generator(device="cpu").manual_seed(7)
    touched = torch.unique(torch.cat([st["src"][:warm], st["dst"][:warm]]))
    model.memory.memory[touched.to(dev)] = (0.1 * torch.randn(touched.numel(), cfg["D"], generator=gen)).to(dev)
    model.memory.last_update[touched.to(dev)] = gs["t"][warm - 1] // 2
    eng.set_cursor(warm)
    return eng, st


def _time_window(eng, steps, world, dev):
    """Device time of `steps` batches between two barriers, max over ranks (ms), plus host issue time per batch."""
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h0 = time.perf_counter()
    eng.run(steps)
    host_ms = (time.perf_counter() - h0) * 1e3 / steps       # host time to ISSUE a batch (no sync inside)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms), host_ms


def run_tgb_eval_bench(args, rank, world, quiet=False, extras=True):
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    steps = args.steps
    E2E = min(steps, 100)                                    # batches of the end-to-end leg
    eng, st = build_eval(rank, world, dev, W + steps + E2E)
    B = eng.B
    eng.run(W)
    torch.cuda.synchronize()
    sampler = None
    if rank == 0:
        from bench import ClockSampler
        sampler = ClockSampler(local)
    ms, host_ms = _time_window(eng, steps, world, dev)
    eng.check()
    mrr = eng.mrr()
    clocks = sampler.stop() if sampler else None
    # ---- end to end: every batch's inputs come from pinned host memory, its metric goes back to the host
    cur = eng.loader.cur_e_id
    sl = slice(cur, cur + E2E * B)
    pin = {k: st[k][sl].contiguous().pin_memory() for k in ("src", "dst", "t", "msg")}
    pin["neg"] = eng.neg[sl].cpu().contiguous().pin_memory()
    acc_host = torch.zeros((2, 2), dtype=torch.float64).pin_memory()
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    h2d = B * (8 * (3 + eng.n_neg) + 4 * eng.Fe)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    seen = 0.0
    for j in range(E2E):
        lo, a, b = cur + j * B, j * B, (j + 1) * B
        eng.src[lo:lo + B].copy_(pin["src"][a:b], non_blocking=True)
        eng.dst[lo:lo + B].copy_(pin["dst"][a:b], non_blocking=True)
        eng.t[lo:lo + B].copy_(pin["t"][a:b], non_blocking=True)
        eng.neg[lo:lo + B].copy_(pin["neg"][a:b], non_blocking=True)
        eng.msg[lo:lo + B].copy_(pin["msg"][a:b], non_blocking=True)
        eng.run(1)
        acc_host[j & 1].copy_(eng.w["acc"], non_blocking=True)       # running (sum of 1/rank, #queries)
        evs[j & 1].record()
        if j > 0:
            evs[(j - 1) & 1].synchronize()
            seen = float(acc_host[(j - 1) & 1][1])                   # read on the host, one batch behind
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": E2E * B / float(dt), "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16,
           "api": "per batch: pinned-host (src,dst,t,neg,msg) -> device stream tables, TGBEvalEngine.run(1), running "
                  "metric -> pinned host (read one batch behind)", "batches": E2E}
    eng.check()
    counts = eng.w["counts"].cpu().tolist()
    out = {"metric": "tgb_eval_queries_per_sec", "value": steps * B / (ms * 1e-3), "unit": "queries/s",
           "n_gpus": world, "steps": steps, "warmup": W, "ms_per_step": ms / steps, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "l2": L2_NOTE},
           "sharding": f"queries of each batch sharded over {world} rank(s), exchange: {eng.exchange_kind}",
           "clocks": clocks, "mrr": mrr, "rank0_subgraph": {"N": counts[0], "E": counts[1]},
           "host_issue_ms_per_step": host_ms, "e2e": e2e, "gpu_launches": steps * eng.launches_per_batch,
           "launches_per_step": eng.launches_per_batch}
    eng.close()
    del eng
    torch.cuda.empty_cache()
    if extras:
        # the 999-negative tgbl-wiki-shaped variant (every rank takes part; extra key of the same line)
        try:
            out["tgbl_wiki_999"] = run_variant(WIKI_TGB, rank, world, dev, min(steps, 60), W)
        except Exception as e:                               # pragma: no cover
            out["tgbl_wiki_999"] = {"unavailable": f"{type(e).__name__}: {e}"}
    if rank == 0 and not quiet:
        print(json.dumps(out))
    if world > 1 and not quiet:
        dist.destroy_process_group()
    return out


def run_variant(cfg, rank, world, dev, steps, W):
    """Device-timed queries/sec of another TGB-eval configuration on the same ranks."""
    eng, st = build_eval(rank, world, dev, W + steps, cfg=cfg)
    eng.run(W)
    torch.cuda.synchronize()
    ms, host_ms = _time_window(eng, steps, world, dev)
    eng.check()
    c = eng.w["counts"].cpu().tolist()
    r = {"metric": "tgb_eval_queries_per_sec", "value": steps * eng.B / (ms * 1e-3), "unit": "queries/s",
         "ms_per_step": ms / steps, "steps": steps, "mrr": eng.mrr(), "rank0_subgraph": {"N": c[0], "E": c[1]},
         "workload": f"tgbl-wiki-shaped: {st['num_nodes']} nodes, 172-d msgs, 200 queries/batch x {cfg['n_neg']} negatives, "
                     f"CTAN_tgb D=T={cfg['D']} S={cfg['S']} K={cfg['K']} (conf_tgb_link_prediction.py:5-16), after a "
                     f"{cfg['warm']}-event warm insert", "launches_per_step": eng.launches_per_batch}
    eng.close()
    del eng
    torch.cuda.empty_cache()
    return r


def run_tgb_eval_reference(args, quiet=False):
    """`--impl reference` for the scaling metric: the reference's OWN `tgb_test(split_mode='val')` (train_link.py:24-203,
    imported unmodified through oracle/ref_loops.py; one model forward per query, numpy MRR per query) on the host
    cores -- same stream / model / negatives / warm state / --steps / --warmup as our arm.  Falls back to the oracle
    port of the same loop when the reference's files are not on the box."""
    from oracle import ctan_oracle as O
    from oracle import ref_loops
    from ctan_b200 import synth
    cfg = COMMENT
    B = cfg["B"]
    W = max(args.warmup, 3)
    n_b = args.steps
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    warm = cfg["warm"]
    n_events = warm + (W + n_b + 4) * B
    st = synth.make_stream("comment", n_events=n_events, seed=0, tgb=True)
    g = torch.Generator().manual_seed(1)
    neg = torch.randint(st["min_dst"], st["max_dst"] + 1, (n_events, cfg["n_neg"]), generator=g)
    mean, std = synth.delta_t_stats(st, 20000)
    kw = dict(num_nodes=st["num_nodes"], edge_dim=st["msg"].size(1), memory_dim=cfg["D"], time_dim=cfg["T"],
              node_dim=1, num_iters=cfg["K"], epsilon=cfg["eps"], gamma=cfg["gamma"], mean_delta_t=mean, std_delta_t=std)
    gen = torch.Generator(device="cpu").manual_seed(7)
    touched = torch.unique(torch.cat([st["src"][:warm], st["dst"][:warm]]))
    mem0 = 0.1 * torch.randn(touched.numel(), cfg["D"], generator=gen)
    lo, hi = warm, warm + (W + n_b) * B
    if ref_loops.available():
        R = ref_loops.load()
        kind = "reference"
        torch.manual_seed(0)
        model = R.ns.CTAN_tgb(**kw)
        model.layernorm = None                                # reference defect (ctdg_models.py:523 reads an attribute it never defines)
        loader = R.ns.LastNeighborLoader(st["num_nodes"], size=cfg["S"])
        for a in range(0, warm, 2000):
            loader.insert(st["src"][a:a + 2000], st["dst"][a:a + 2000])
        model.memory.memory[touched] = mem0
        model.memory.last_update[touched] = st["t"][warm - 1] // 2
        helper = torch.empty(st["num_nodes"], dtype=torch.long)
        data = R.TemporalData(src=st["src"], dst=st["dst"], t=st["t"], msg=st["msg"])
        win = R.TemporalData(src=st["src"][lo:hi], dst=st["dst"][lo:hi], t=st["t"][lo:hi], msg=st["msg"][lo:hi])
        win.x = st["x"]
        tl = ref_loops.TimedLoader(R.TemporalDataLoader(win, batch_size=B), skip=W)
        ev = R.Evaluator(name="tgbl-comment")
        res, _ = R.train_link.tgb_test(data, model, tl, ref_loops.FixedEvalNegatives(neg, start=lo), loader, "val",
                                       helper, ev, "mrr", device="cpu")
        dt, mrr = tl.seconds, res["mrr"]                      # (the reference's mean also covers the warm-up batches)
    else:
        kind = "port"
        torch.manual_seed(0)
        om = O.OracleCTAN(tgb=True, **kw)
        loader = O.TorchNeighborLoader(st["num_nodes"], cfg["S"])
        for a in range(0, warm, 2000):
            loader.insert(st["src"][a:a + 2000], st["dst"][a:a + 2000])
        om.memory.memory[touched] = mem0
        om.memory.last_update[touched] = st["t"][warm - 1] // 2
        helper = torch.zeros(st["num_nodes"], dtype=torch.long)
        mrrs = []
        with torch.no_grad():
            for it in range(W + n_b):
                if it == W:
                    t0 = time.perf_counter()
                a, b = lo + it * B, lo + (it + 1) * B
                m, _, _ = O.tgb_eval_batch_ref(om, loader, helper, st, a, b, [neg[q].tolist() for q in range(a, b)])
                mrrs += m
            dt = time.perf_counter() - t0
        mrr = float(np.mean(mrrs))
    v = n_b * B / dt
    what = ("the reference's own tgb_test('val') (train_link.py:24-203), unmodified" if kind == "reference"
            else "oracle port of train_link.py:111-201")
    out = {"impl": "reference", "metric": "tgb_eval_queries_per_sec", "value": v, "unit": "queries/s",
           "n_gpus": args.gpus, "steps": n_b, "warmup": W, "ms_per_step": dt / n_b * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "l2": L2_NOTE},
           "mrr": mrr,
           "cpu_baseline": {"value": v, "unit": "queries/s", "cores": cores, "kind": kind,
                            "sample": f"{n_b} consecutive 200-query batches after {W} warm-up batches ({what}; one model "
                                      "forward per query on the host cores)"},
           "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not quiet:
        print(json.dumps(out))
    return out
