#!/usr/bin/env python
"""Benchmark of the CTAN per-event-batch hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1  : train events/sec of the full per-batch step (lookup -> forward -> loss -> update+insert ->
         backward -> Adam) on the Wikipedia-shaped synthetic stream (BASELINE.json configs[1]); a "step"
         is one 200-event batch.  Extra keys carry infer events/sec, the single-GPU TGB-eval queries/sec,
         the eager-PyTorch-on-GPU baseline and the CPU baseline.
N > 1  : TGB one-vs-many evaluation (tgb_test val/test semantics, train_link.py:111-199) with the
         queries of every 200-event batch sharded over N ranks and ONE NCCL all-gather per batch
         (launched by torchrun, one rank per GPU); value = queries/sec of the whole job.
--impl reference : the reference algorithm's CPU implementation (oracle port: /root/reference is pure
         Python and cannot travel to the GPU box) timed on the host cores for the same workload.
Prints ONE JSON line on rank 0."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WIKI = dict(D=128, T=16, S=5, H=64, B=200, Fe=172, eps=0.1, gamma=0.1, lr=1e-3)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured sustained bf16 (MEASURED_PEAKS.json)"
    return 1400.0, "fallback sustained (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if "Active" in v and "Not" not in v:
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def make_workload(n_events, seed=0, shape="wiki"):
    import ctan_b200  # noqa: F401
    from ctan_b200 import synth
    st = synth.make_stream(shape, n_events=n_events, seed=seed)
    st["neg"] = synth.draw_negatives(st)[:, 0]
    n_train = int(0.7 * st["src"].numel())
    mean, std = synth.delta_t_stats(st, min(n_train, 20000))
    return st, mean, std


def model_kwargs(st, K, mean, std, cfg=WIKI):
    return dict(num_nodes=st["num_nodes"], edge_dim=st["msg"].size(1), memory_dim=cfg["D"], time_dim=cfg["T"],
                node_dim=1, num_iters=K, epsilon=cfg["eps"], gamma=cfg["gamma"], readout_hidden=cfg["H"],
                mean_delta_t=mean, std_delta_t=std, final_act=None)


# ------------------------------------------------------------------------------------------ baselines
def reference_train_rate(st, K, mean, std, device, n_steps, warmup, threads=None, ours=False, cfg=None):
    """events/sec of the reference's OWN `train_link.train` (train_link.py:287-345), imported UNMODIFIED through
    oracle/ref_loops.py (PyG operators from oracle/pyg_shim.py; the files come from /root/reference or its git-ignored
    mirror baseline/_ref/), on `device`, fed the same pre-drawn negatives as every other implementation.  The loop
    function is called once on W+K batches; a timestamping loader wrapper isolates the last K.
    ours=True: the same call with ctan_b200.models swapped in for the reference's classes -- what a user gets from
    the two-line import change of INTEGRATION.md.  Returns (events/s, ms/step, kind)."""
    from oracle import ref_loops
    cfg = cfg or WIKI
    B = cfg["B"]
    if not ref_loops.available():
        if ours:
            raise RuntimeError("baseline/_ref/ is missing: run `python __graft_entry__.py build` where /root/reference exists")
        v, ms = oracle_train_rate(st, K, mean, std, device, n_steps, warmup, threads, cfg)
        return v, ms, "port"
    R = ref_loops.load()
    if threads:
        torch.set_num_threads(threads)
    n = (warmup + n_steps) * B
    data = R.TemporalData(src=st["src"][:n].to(device), dst=st["dst"][:n].to(device), t=st["t"][:n].to(device),
                          msg=st["msg"][:n].to(device))
    data.x = st["x"].to(device)
    torch.manual_seed(0)
    if ours:
        from ctan_b200.models import CTAN, LastNeighborLoader
        model = CTAN(**model_kwargs(st, K, mean, std, cfg)).to(device)
        loader = LastNeighborLoader(st["num_nodes"], cfg["S"], device=device)
    else:
        model = R.ns.CTAN(**model_kwargs(st, K, mean, std, cfg)).to(device)
        loader = R.ns.LastNeighborLoader(st["num_nodes"], size=cfg["S"], device=device)
    helper = torch.empty(st["num_nodes"], dtype=torch.long, device=device)
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    sync = torch.cuda.synchronize if str(device) != "cpu" else None
    tl = ref_loops.TimedLoader(R.TemporalDataLoader(data, batch_size=B), skip=warmup, sync=sync)
    R.train_link.train(data, model, opt, tl, torch.nn.BCEWithLogitsLoss(), loader, helper,
                       train_neg_sampler=ref_loops.FixedNegatives(st["neg"].view(-1)), device=device)
    return tl.n_timed * B / tl.seconds, tl.seconds / tl.n_timed * 1e3, "reference"


def oracle_train_rate(st, K, mean, std, device, n_steps, warmup, threads=None, cfg=None):
    """events/sec of the restated reference training loop (oracle port; used only when the reference's own files are
    not available on the box) on `device`."""
    from oracle import ctan_oracle as O
    if threads:
        torch.set_num_threads(threads)
    cfg = cfg or WIKI
    torch.manual_seed(0)
    om = O.OracleCTAN(**model_kwargs(st, K, mean, std, cfg)).to(device)
    ld = O.TorchNeighborLoader(st["num_nodes"], cfg["S"], device=device)
    h = torch.zeros(st["num_nodes"], dtype=torch.long, device=device)
    opt = torch.optim.Adam(om.parameters(), lr=cfg["lr"])
    g = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in st.items()}
    B = cfg["B"]

    def sync():
        if device != "cpu":
            torch.cuda.synchronize()
    for it in range(warmup):
        O.train_step_ref(om, ld, h, opt, g, it * B, (it + 1) * B, g["neg"][it * B:(it + 1) * B])
    sync()
    t0 = time.perf_counter()
    for it in range(warmup, warmup + n_steps):
        O.train_step_ref(om, ld, h, opt, g, it * B, (it + 1) * B, g["neg"][it * B:(it + 1) * B])
    sync()
    dt = time.perf_counter() - t0
    return n_steps * B / dt, dt / n_steps * 1e3


# ------------------------------------------------------------------------------------------ ours, N=1
def build_engine(st, K, mean, std, dev, host_staged=False, cfg=None):
    import ctan_b200  # noqa: F401
    from ctan_b200.engine import StepEngine
    from ctan_b200.models import CTAN, LastNeighborLoader
    cfg = cfg or WIKI
    torch.manual_seed(0)
    model = CTAN(**model_kwargs(st, K, mean, std)).to(dev)
    loader = LastNeighborLoader(st["num_nodes"], cfg["S"], device=dev)
    if host_staged:
        g = dict(st)                                           # stays on the host; batches are pushed per step
    else:
        g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in st.items()}
        g["neg"] = st["neg"].view(-1, 1).to(dev)
    eng = StepEngine(model, loader, g, batch_size=cfg["B"], n_neg=1, lr=cfg["lr"], host_staged=host_staged)
    eng.reset(reset_optimizer=True)
    eng.warmup()
    return eng


def time_steps(eng, n_steps, train, flush_buf):
    """Per-step CUDA events with an L2 flush (write of a >L2 buffer) between timed steps."""
    evs = []
    for fn in eng.plan(n_steps, train):      # the engine's own replay sequence (a training step also issues the
        if flush_buf is not None:            # next step's front-end beside its backward: that work is timed here)
            flush_buf.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    eng.loader.cur_e_id += n_steps * eng.B
    return [a.elapsed_time(b) for a, b in evs]


def time_back_to_back(eng, n_steps, train):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.run(n_steps, train=train)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def stage_timeline(eng, n_replays, train=True):
    """Live per-stage device durations INSIDE the CUDA-graph replay: a one-thread probe kernel storing
    %globaltimer is captured after every C-ABI call of the critical-path stream (side-stream work overlaps
    and shows up only where the critical path has to join it).  Returns ({call: us}, launches_per_step)."""
    from ctan_b200 import _lib
    ts = torch.zeros(256, dtype=torch.int64, device=eng.dev)
    names = []
    g = torch.cuda.CUDAGraph()
    _lib.N_LAUNCHES = 0
    with torch.cuda.graph(g):
        _lib.lib().ctan_probe(ts.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        names.append("start")
        _lib.PROBE = (ts, names, torch.cuda.current_stream().cuda_stream)
        eng._launch(train)
        _lib.PROBE = None
    launches = _lib.N_LAUNCHES
    acc = torch.zeros(len(names), dtype=torch.float64)
    for _ in range(n_replays):
        g.replay()
        torch.cuda.synchronize()
        t = ts[:len(names)].cpu().double()
        acc += t - t[0]
    eng.loader.cur_e_id += n_replays * eng.B
    acc = (acc / n_replays / 1e3).tolist()
    out = {}
    for i in range(1, len(names)):
        out[names[i]] = out.get(names[i], 0.0) + max(acc[i] - acc[i - 1] - 1.3, 0.1)   # ~1.3 us is the probe itself
    return out, launches


def saturating_edge_roofline(dev, N=100_000, S=16, D=256, iters=10):
    """The fused edge kernel (logits + segmented softmax + aggregation + tanh update) on a shape that
    saturates the memory system: N rows of degree S gathering from a node table larger than L2.
    Algorithmic bytes: 16*D per node (q, xA+b, x_in, x_out) + (12*D+8) per edge (k_j, v_j, EP row, src id)."""
    from ctan_b200._lib import call, ptr
    E = N * S
    g = torch.Generator(device="cpu").manual_seed(0)
    QKVA = torch.randn(N, 4 * D, device=dev)
    EP = torch.randn(E, D, device=dev)
    X = torch.randn(N, D, device=dev)
    Xo = torch.empty(N, D, device=dev)
    rowptr = (torch.arange(N + 1, dtype=torch.int32) * S).to(dev)
    esrc = torch.randint(0, N, (E,), generator=g).to(dev)

    def fn():
        call("ctan_edge_fwd", ptr(QKVA), ptr(EP), ptr(rowptr), ptr(esrc), N, None, D, ptr(X), 0.5, ptr(Xo), None, None,
             None, None, 0, 0)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    alg = N * 16 * D + E * (12 * D + 8)
    del QKVA, EP, X, Xo
    torch.cuda.empty_cache()
    return {"kernel": "edge_fwd_kernel", "shape": {"N": N, "E": E, "D": D}, "ms_per_launch": ms,
            "algorithmic_bytes": alg, "achieved": alg / (ms * 1e-3) / 1e9}


def run_e2e(st, K, mean, std, dev, n_steps, warmup):
    """Same step driven with HOST buffers through the public engine API: every step copies the batch
    (src,dst,t,neg ids + msg rows) from pinned host memory, runs the step, and reads the loss back."""
    eng = build_engine(st, K, mean, std, dev, host_staged=True)
    B = eng.B
    pin = {k: st[k].pin_memory() for k in ("src", "dst", "t", "msg")}
    pin["neg"] = st["neg"].view(-1, 1).pin_memory()
    loss_host = torch.zeros(2).pin_memory()                  # two pinned slots: step i's loss is read while i+1 runs
    done = [torch.cuda.Event(), torch.cuda.Event()]
    h2d = 0

    nb = st["src"].numel() // B
    ids_pin = torch.cat([st["src"][:nb * B].view(nb, B), st["dst"][:nb * B].view(nb, B), st["t"][:nb * B].view(nb, B),
                         st["neg"][:nb * B].view(nb, B)], 1).contiguous().pin_memory()   # the loader's packed batches

    msg_rows = pin["msg"][:nb * B].view(nb, B, -1)           # (views made once: the loop below is host-bound)
    loss_dst = [loss_host[0:1], loss_host[1:2]]
    loss_src = [eng.w["loss_hist"][i:i + 1] for i in range(min(eng.hist_cap, warmup + n_steps + 1))]

    def issue(it):
        nonlocal h2d
        h2d = eng.push_packed(ids_pin[it], msg_rows[it])
        eng.run(1, train=True)
        loss_dst[it & 1].copy_(loss_src[it % len(loss_src)] if len(loss_src) == eng.hist_cap else loss_src[it], non_blocking=True)
        done[it & 1].record()

    def collect(it):                                         # host read of step `it`'s loss (blocks until it is there)
        done[it & 1].synchronize()
        return float(loss_host[it & 1])
    for it in range(warmup):
        issue(it)
        collect(it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    last = 0.0
    for it in range(warmup, warmup + n_steps):
        issue(it)                                            # H2D of this step's batch + the step + D2H of its loss
        if it > warmup:
            last = collect(it - 1)                           # every step's loss reaches the host, one step behind
    last = collect(warmup + n_steps - 1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    eng.check()
    return n_steps * B / dt, h2d, 4, last


def other_configs(dev, W, flush):
    """Short engine runs on the other single-GPU BASELINE.json shapes: C3 (Reddit-shaped, 10 984 nodes / 672 447
    events) and C2 with K=3.  Same per-step CUDA-event timing with the L2 flush."""
    res = {}
    B = WIKI["B"]
    for name, shape, K in (("reddit_K1", "reddit", 1), ("wiki_K3", "wiki", 3), ("wiki_K5", "wiki", 5)):
        n_steps = 100
        st, mean, std = make_workload(None if shape == "reddit" else 157474, shape=shape)
        eng = build_engine(st, K, mean, std, dev)
        eng.run(W, train=True)
        torch.cuda.synchronize()
        tr = sum(time_steps(eng, n_steps, True, flush))
        eng.check()
        eng.reset()
        eng.run(W, train=False)
        inf = sum(time_steps(eng, n_steps, False, flush))
        eng.check()
        c = eng.w["counts"].cpu().tolist()
        res[name] = {"train_events_per_sec": n_steps * B / (tr * 1e-3), "infer_events_per_sec": n_steps * B / (inf * 1e-3),
                     "train_ms_per_step": tr / n_steps, "steps": n_steps, "N": c[0], "E": c[1],
                     "events_in_stream": int(st["src"].numel()), "nodes": int(st["num_nodes"])}
        del eng, st
        torch.cuda.empty_cache()
    res["pascal_shaped_multiclass_K3"] = pascal_shaped(dev, W, flush)
    res["path_graph_sequence_L20_K3"] = path_graph_sequence(dev)
    return res


def pascal_shaped(dev, W, flush, n_steps=100):
    """BASELINE.json configs[3] (C4, Temporal Pascal-VOC shape: conf_pascal.py:136-175, run_pascal.py:31): 256-event
    batches over 1e5 nodes, 1-d messages, 14-d node features, 21-class head with CrossEntropy and NO negatives, CTAN D=74
    T=16 S=5 K=3 H=37 -- the engine's labelled head.  Same timing rule as the headline (per-step events, L2 flushed)."""
    import ctan_b200  # noqa: F401
    from ctan_b200 import synth
    from ctan_b200.engine import StepEngine
    from ctan_b200.models import CTAN, LastNeighborLoader
    B, K = 256, 3
    st = synth.make_stream("comment", n_events=(W + n_steps + 8) * B, seed=0, msg_dim=1, n_nodes_scale=100_000 / 994_790)
    g0 = torch.Generator().manual_seed(0)
    x = torch.randn(st["num_nodes"], 14, generator=g0)
    y = torch.randint(0, 21, (st["src"].numel(),), generator=g0)
    mean, std = synth.delta_t_stats(st, 20000)
    torch.manual_seed(0)
    model = CTAN(num_nodes=st["num_nodes"], edge_dim=1, memory_dim=74, time_dim=16, node_dim=14, num_iters=K, epsilon=0.1,
                 gamma=0.1, readout_hidden=37, readout_out=21, mean_delta_t=mean, std_delta_t=std).to(dev)
    loader = LastNeighborLoader(st["num_nodes"], 5, device=dev)
    g = {k: st[k].to(dev) for k in ("src", "dst", "t", "msg")}
    g["x"], g["y"] = x.to(dev), y.to(dev)
    eng = StepEngine(model, loader, g, batch_size=B, n_neg=0, lr=1e-4, weight_decay=1e-6, criterion=torch.nn.CrossEntropyLoss())
    eng.reset(reset_optimizer=True)
    eng.warmup()
    eng.run(W, train=True)
    torch.cuda.synchronize()
    tr = sum(time_steps(eng, n_steps, True, flush))
    eng.check()
    c = eng.w["counts"].cpu().tolist()
    return {"train_events_per_sec": n_steps * B / (tr * 1e-3), "train_ms_per_step": tr / n_steps, "steps": n_steps, "batch": B,
            "N": c[0], "E": c[1], "nodes": int(st["num_nodes"]), "last_loss": float(eng.losses(1)[0]),
            "contractions": "CUDA-core fp32 (D=74 is not a multiple of 32)"}


def path_graph_sequence(dev, n_seq=200, L=20):
    """BASELINE.json configs[0] (C1, Temporal Path Graph: datasets/label_prediction_sequence.py:27-57, conf_sequence.py:
    134-170): sequences of L-1 one-event steps, SequencePredictor head, CTAN D=53 K=3, meta-batches of 16 sequences, through
    ctan_b200.train_sequence.train (the reference's signature; engine sequence mode).  Wall clock of one whole pass."""
    import ctan_b200  # noqa: F401
    from ctan_b200 import train_link as TL
    from ctan_b200 import train_sequence as TS
    from ctan_b200.models import CTAN, LastNeighborLoader
    g0 = torch.Generator().manual_seed(0)
    num_nodes = n_seq * L
    x = (torch.rand(1, num_nodes, 1, generator=g0) * 2 - 1).to(dev)
    data = []
    for i in range(n_seq):
        s = torch.arange(i * L, i * L + L - 1)
        data.append(TL.TemporalData(src=s.to(dev), dst=(s + 1).to(dev), t=torch.arange(L - 1).to(dev),
                                    msg=(torch.rand(L - 1, 1, generator=g0) * 2 - 1).to(dev),
                                    y=torch.tensor([float(i % 2)]).to(dev), x=x))
    torch.manual_seed(0)
    model = CTAN(num_nodes=num_nodes, edge_dim=1, memory_dim=53, time_dim=1, node_dim=1, num_iters=3, epsilon=0.5, gamma=0.1,
                 readout_hidden=26, readout_out=1, mean_delta_t=1.0, std_delta_t=1.0, final_act="tanh", predict_dst=True).to(dev)
    loader = LastNeighborLoader(num_nodes, 5, device=dev)
    helper = torch.empty(num_nodes, dtype=torch.long, device=dev)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3, weight_decay=1e-5)
    crit = torch.nn.BCEWithLogitsLoss()
    meta = [torch.arange(i, min(i + 16, n_seq)) for i in range(0, n_seq, 16)]
    loaders = [TL.TemporalDataLoader(d, batch_size=1) for d in data]
    TS.train(data, model, opt, meta, loaders, crit, loader, helper, device=dev)          # builds + warms the engine
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    TS.train(data, model, opt, meta, loaders, crit, loader, helper, device=dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n_ev = n_seq * (L - 1)
    return {"train_events_per_sec": n_ev / dt, "us_per_event": dt / n_ev * 1e6, "sequences": n_seq, "events": n_ev,
            "what": "wall clock of one ctan_b200.train_sequence.train pass (one graph replay per event, optimizer step "
                    "per 16 sequences)"}


def dropin_engine_rate(st, K, mean, std, dev, n_steps):
    """events/sec of `ctan_b200.train_link.train` -- the reference's loop SIGNATURE (train_link.py:287) on the engine:
    same arguments, same return value (per-batch losses), one call."""
    from ctan_b200 import train_link as TL
    from ctan_b200.models import CTAN, LastNeighborLoader
    from oracle import ref_loops
    cfg = WIKI
    B = cfg["B"]
    n = n_steps * B
    torch.manual_seed(0)
    model = CTAN(**model_kwargs(st, K, mean, std)).to(dev)
    loader = LastNeighborLoader(st["num_nodes"], cfg["S"], device=dev)
    helper = torch.empty(st["num_nodes"], dtype=torch.long, device=dev)
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    data = TL.TemporalData(src=st["src"][:n].to(dev), dst=st["dst"][:n].to(dev), t=st["t"][:n].to(dev),
                           msg=st["msg"][:n].to(dev))
    data.x = st["x"].to(dev)
    tl = TL.TemporalDataLoader(data, batch_size=B)
    sampler = ref_loops.FixedNegatives(st["neg"].view(-1))
    crit = torch.nn.BCEWithLogitsLoss()
    TL.train(data, model, opt, tl, crit, loader, helper, train_neg_sampler=sampler, device=dev)   # builds + warms the engine
    sampler.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    losses = TL.train(data, model, opt, tl, crit, loader, helper, train_neg_sampler=sampler, device=dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "events/s", "ms_per_step": dt / n_steps * 1e3, "batches": n_steps,
            "last_loss": float(losses[-1]),
            "what": "ctan_b200.train_link.train(data, model, optimizer, train_loader, criterion, neighbor_loader, helper, "
                    "train_neg_sampler, device): the reference's signature and return value, StepEngine underneath; "
                    "wall clock of one whole call (one epoch of the sample), negatives drawn by the sampler per batch"}


def pick_metric(args, world):
    """Which of BASELINE.json's two metrics is this run's `value`?
    * train events/sec (configs[1], one B200): the headline of a single-GPU run -- the training step does not shard.
    * TGB-eval queries/sec (configs[4]): the metric that scales over GPUs.  It is the `value` of every run of a SCALING
      series -- N > 1, any run under torchrun, and also the N=1 point taken on a multi-GPU box -- so that the 1/2/4/8
      curve is ONE metric; each line carries the other metric as an extra key.  `--metric` overrides."""
    if args.metric != "auto":
        return args.metric
    if world > 1 or args.gpus > 1 or "WORLD_SIZE" in os.environ:
        return "tgb_eval"
    try:
        if torch.cuda.device_count() > 1 or _visible_gpus() > 1:
            return "tgb_eval"
    except Exception:
        pass
    return "train"


def _visible_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=1, help="CTAN num_iters (K)")
    ap.add_argument("--metric", default="auto", choices=["auto", "train", "tgb_eval"])
    ap.add_argument("--no-baselines", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(args.warmup, 3)
    K = args.iters
    B = WIKI["B"]
    metric = pick_metric(args, world)
    cfg_desc = {"workload": "wiki-shaped synthetic link prediction: 9227 nodes, 157474 events, 172-d msgs, batch 200; "
                            f"CTAN D=128 T=16 S=5 K={K} H=64, 1 negative/positive, full train step incl. Adam",
                "l2": "flushed between timed steps (256 MiB write), per-step CUDA events summed",
                "metric_rule": "single-GPU box: train events/s (this line); scaling series (multi-GPU box, any N): "
                               "TGB-eval queries/s, so the N=1 point of a scaling run differs from this line BY DESIGN"}

    if args.impl == "reference":
        if rank != 0:
            return
        if metric == "tgb_eval":                        # the scaling metric is TGB evaluation: time the same on the CPU
            from bench_tgb import run_tgb_eval_reference
            run_tgb_eval_reference(args)
            return
        n_steps = args.steps
        st, mean, std = make_workload(max(157474, (W + n_steps + 64) * B))
        if st["src"].numel() != 157474:
            cfg_desc["workload"] += f" (stream lengthened to {st['src'].numel()} events to fit --steps)"
        cores = os.cpu_count() or 1
        v, ms, kind = reference_train_rate(st, K, mean, std, "cpu", n_steps, W, threads=cores)
        what = ("the reference's own train_link.train (train_link.py:287-345) imported unmodified, PyG operators from "
                "oracle/pyg_shim.py" if kind == "reference" else "oracle port of train_link.py:287-345")
        print(json.dumps({"impl": "reference", "metric": "train_events_per_sec", "value": v, "unit": "events/s",
                          "n_gpus": args.gpus, "steps": n_steps, "warmup": W, "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": cfg_desc,
                          "cpu_baseline": {"value": v, "unit": "events/s", "cores": cores, "kind": kind,
                                           "sample": f"{n_steps} consecutive 200-event training batches of the same "
                                                     f"stream after {W} warm-up batches ({what})"},
                          "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if metric == "tgb_eval":
        from bench_tgb import run_tgb_eval_bench        # scaling series: sharded TGB evaluation
        run_tgb_eval_bench(args, rank, world, extras=not args.no_baselines)
        return

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    from ctan_b200 import _lib
    n_events = max(157474, (W + args.steps + 64) * B)          # the stream is lengthened only if K asks for more
    if n_events != 157474:
        cfg_desc["workload"] += f" (stream lengthened to {n_events} events to fit --steps)"
    st, mean, std = make_workload(n_events)
    eng = build_engine(st, K, mean, std, dev)
    flush = torch.zeros(64 * 1024 * 1024, device=dev)           # 256 MiB > 126 MB L2
    # ---- phase A: warm-up, then K timed training steps (L2 flushed between steps)
    eng.run(W, train=True)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    ms_list = time_steps(eng, args.steps, True, flush)
    clocks = sampler.stop()
    total_ms = sum(ms_list)
    value = args.steps * B / (total_ms * 1e-3)
    eng.check()
    losses = eng.losses(5).cpu().tolist()
    # ---- phase B: per-stage device timeline inside the graph replay (probe kernels on the critical path)
    prof, launches_per_step = stage_timeline(eng, 50, train=True)
    counts = eng.w["counts"].cpu().tolist()
    # ---- phase C: same stream from a fresh state, back-to-back replays (L2 warm, as in production)
    eng.reset(reset_optimizer=True)
    eng.run(W, train=True)
    warm_ms = time_back_to_back(eng, args.steps, True)
    warm_value = args.steps * B / (warm_ms * 1e-3)
    # ---- phase D: inference (no-grad) steps from a fresh state
    eng.reset()
    eng.run(W, train=False)
    inf_ms = sum(time_steps(eng, args.steps, False, flush))
    infer_value = args.steps * B / (inf_ms * 1e-3)
    eng.check()

    # ---- roofline of the dominant kernel (by live per-call device time)
    peak, peak_src = load_peaks()
    N_last, E_last = counts[0], counts[1]
    top = sorted(prof.items(), key=lambda kv: -kv[1])
    top_name, top_us = top[0]
    D, Fe, T, Hd = eng.D, eng.Fe, eng.T, eng.Hd
    alg_bytes = {
        "ctan_eproj_fwd": E_last * (4 * Fe + 8 + 4) + 4 * D * (Fe + T) + E_last * D * 4,
        "ctan_eproj_bwd": E_last * (4 * Fe + 8 + 4 + 4 * D) + 2 * 4 * D * (Fe + T) + E_last * T * 4,
        "ctan_qkva_fwd": N_last * D * 4 + 4 * D * D * 4 + N_last * 4 * D * 4,
        "ctan_qkva_fwd_tc": N_last * D * 4 + 2 * 4 * D * D * 4 + N_last * 4 * D * 4,
        "ctan_dx_tc": N_last * (4 * D + D + D) * 4 + 2 * 4 * D * D * 4,
        "ctan_qkva_bwd": N_last * (4 * D + D + D) * 4 + 2 * 4 * D * D * 4 + N_last * D * 4,
        "ctan_edge_fwd": N_last * (4 * D + 2 * D) * 4 + E_last * (2 * D * 4 + D * 4 + 8 + 4),
        "ctan_edge_bwd": N_last * (4 * D + 3 * D + 4 * D) * 4 + E_last * (3 * D * 4 + 2 * D * 4 + 8 + 8),
        "ctan_readout_fwd": eng.P * (2 * D + Hd + 1) * 4 + 2 * Hd * D * 4,
        "ctan_readout_bwd": eng.P * (4 * D + 3 * Hd + 2) * 4 + 4 * Hd * D * 4,
        "ctan_enc_fwd": N_last * (2 * D + 1) * 4 + D * (D + 1) * 4,
        "ctan_enc_bwd": N_last * (2 * D + 1) * 4 + D * (D + 1) * 4,
        # folded enc_x + projection: gathered rows [N, Kx] + hi/lo panel [5D, Kx] in, X0 [N, D] + QKVA [N, 4D] out
        "ctan_enc_qkva_tc": N_last * getattr(eng, "Kx", D + 1) * 4 + 2 * 5 * D * getattr(eng, "Kx", D + 1) * 4
                            + N_last * 5 * D * 4,
        "ctan_enc_prep_ext": N_last * (D + eng.Fn) * 4 + N_last * 8 + N_last * getattr(eng, "Kx", D + 1) * 4,
        # one-launch link readout: two embedding rows per pair + both weights in; H, dHid, OUT, dOUT out; dZ atomics
        "ctan_readout_link_fused": eng.P * (2 * D * 4 + 16) + 2 * Hd * D * 4 + eng.P * (2 * Hd + 2) * 4
                                   + eng.P * 2 * D * 4,
    }.get(top_name)
    roof = {"bound": "hbm", "kernel": top_name, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
            "traffic": None, "peak_source": peak_src, "ms_per_launch": top_us * 1e-3,
            "how": "longest stage of the in-graph critical-path timeline (globaltimer probes, 50 replays)",
            "note": "per-batch work is ~3 MB / ~0.2 GFLOP: every kernel of the step is latency-bound at B=200 "
                    "(DESIGN.md S4); the saturating-shape figure is under roofline_saturating"}
    if alg_bytes:
        ach = alg_bytes / (top_us * 1e-6) / 1e9
        roof.update(achieved=ach, frac=ach / peak, algorithmic_bytes=alg_bytes)
    # a tcgen05 contraction is bounded by the tensor pipe, not HBM: report useful flops against the measured dense peak
    tc_flops = {"ctan_dx_tc": 2.0 * N_last * 4 * D * D, "ctan_qkva_fwd_tc": 2.0 * N_last * D * 4 * D,
                "ctan_enc_qkva_tc": 2.0 * N_last * (D + eng.Fn) * 5 * D}.get(top_name)
    # roofline.traffic = DRAM bytes of ONE launch of this kernel at this shape from the committed `ncu --set full`
    # capture of the round (DRAM counters need the profiler: they are not measured in this run, and the field says
    # where they come from): profiles/r02_traffic.json, keyed by C-ABI entry point (+ the two older tcgen05 keys)
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if K == 1 and os.path.exists(tp):
        tj = json.load(open(tp))
        tr = tj.get(top_name + "@wiki") or tj.get({"ctan_dx_tc": "qkva_tc_kernel@wiki_dx",
                                                   "ctan_qkva_fwd_tc": "qkva_tc_kernel@wiki_qkva"}.get(top_name, ""))
        if tr:
            roof["traffic"] = tr["traffic_bytes"]
            roof["traffic_source"] = tr.get("source", "ncu --set full capture under profiles/")
    if tc_flops:
        tpeak, tsrc = load_tensor_peak()
        ach = tc_flops / (top_us * 1e-6) / 1e12
        roof.update(bound="tensor", unit="TFLOP/s", achieved=ach, peak=tpeak, frac=ach / tpeak, peak_source=tsrc,
                    algorithmic_flops=tc_flops, hbm_view={"algorithmic_bytes": alg_bytes,
                                                          "achieved_gbs": alg_bytes / (top_us * 1e-6) / 1e9},
                    precision="3xTF32 (three tf32 MMAs per useful product; the peak is the measured dense bf16 figure, "
                              "tf32 runs at half of it)")
    # whole-step algorithmic bytes (SURVEY.md S8d formula)
    U = 2 * B
    Pw = eng.n_params
    step_bytes = (16 * eng.S * N_last + N_last * (4 * D + 8 + 4) + E_last * (4 * Fe + 8) + 4 * Pw + 4 * eng.P
                  + U * (4 * D + 8) + 2 * 16 * eng.S * U + 24 * B + 8 * Pw)
    step_ms = total_ms / args.steps
    roof["step"] = {"algorithmic_bytes": step_bytes, "achieved_gbs": step_bytes / (step_ms * 1e-3) / 1e9,
                    "frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak, "N": N_last, "E": E_last}

    out = {"metric": "train_events_per_sec", "value": value, "unit": "events/s", "n_gpus": 1, "steps": args.steps,
           "warmup": W, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": cfg_desc, "clocks": clocks,
           "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
           "value_l2_warm_back_to_back": warm_value, "infer_events_per_sec": infer_value,
           "last_losses": losses, "roofline": roof,
           "critical_path_us": {k: round(v, 2) for k, v in top}}
    sat = saturating_edge_roofline(dev)
    sat.update(bound="hbm", peak=peak, unit="GB/s", frac=sat["achieved"] / peak, traffic=None)
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        tr = json.load(open(tpath)).get("edge_fwd_kernel@saturating(N=100000,E=1600000,D=256)")
        if tr:
            sat["traffic"] = tr["traffic_bytes"]
    out["roofline_saturating"] = sat
    # the scatter side (edge backward: atomic accumulation into source rows) at the same shape
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import sat_edge_bwd
    sb = sat_edge_bwd.run(dev=dev)
    sb = {"kernel": sb["kernel"], "shape": {"N": sb["N"], "E": sb["E"], "D": sb["D"]}, "ms_per_launch": sb["ms_per_launch"],
          "algorithmic_bytes": sb["algorithmic_bytes"], "achieved": sb["achieved_gbs"], "bound": "hbm", "peak": peak,
          "unit": "GB/s", "frac": sb["achieved_gbs"] / peak, "traffic": None,
          "note": "atomics counted as read-modify-write (16*D bytes per edge)"}
    if os.path.exists(tpath):
        tr = json.load(open(tpath)).get("edge_bwd_kernel@saturating(N=100000,E=1600000,D=256)")
        if tr:
            sb["traffic"] = tr["traffic_bytes"]
    out["roofline_saturating_scatter"] = sb
    torch.cuda.empty_cache()

    # ---- end to end with host buffers
    e2e_v, h2d, d2h, last_loss = run_e2e(st, K, mean, std, dev, min(args.steps, 400), W)
    out["e2e"] = {"value": e2e_v, "unit": "events/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "api": "StepEngine(host_staged=True): per batch push_packed(pinned ids, pinned msg rows) -> device over a "
                         "copy stream (double-buffered staging), run(1), loss -> pinned host; the host reads step i's "
                         "loss while step i+1 runs (every loss is read, one step behind)"}

    if not args.no_baselines:
        del eng
        torch.cuda.empty_cache()
        # ---- other BASELINE.json shapes of the same path, short runs (extra keys; same engine, same timing rule)
        out["other_configs"] = other_configs(dev, W, flush)
        # ---- the GPU-PyTorch baseline north_star's ">= 20x" is quoted against: the reference's own loop on this B200
        gst, gmean, gstd = (st, mean, std) if st["src"].numel() >= 520 * B else make_workload(520 * B)
        gv, gms, gkind = reference_train_rate(gst, K, gmean, gstd, "cuda:0", 500, 20)
        out["torch_gpu_baseline"] = {"value": gv, "unit": "events/s", "ms_per_step": gms, "kind": gkind, "batches": 500,
                                     "what": "the reference's own train_link.train (unmodified, PyG operators from the "
                                             "shim) in eager PyTorch on the same B200, fp32, 500 batches after 20"
                                             if gkind == "reference" else "oracle port, eager PyTorch on the same B200"}
        out["speedup_vs_torch_gpu"] = {"device_timed": value / gv, "e2e": e2e_v / gv}
        try:
            dv, dms, _ = reference_train_rate(gst, K, gmean, gstd, "cuda:0", 300, 20, ours=True)
            out["drop_in_api"] = {"value": dv, "unit": "events/s", "ms_per_step": dms,
                                  "what": "the reference's own train_link.train, UNMODIFIED, on ctan_b200.models (CTAN + "
                                          "LastNeighborLoader swapped in, torch.optim.Adam, loss.item() per step)"}
        except RuntimeError as e:
            out["drop_in_api"] = {"unavailable": str(e)}
        try:
            out["drop_in_engine_loop"] = dropin_engine_rate(gst, K, gmean, gstd, dev, 300)
        except Exception as e:                                   # pragma: no cover
            out["drop_in_engine_loop"] = {"unavailable": f"{type(e).__name__}: {e}"}
        cores = os.cpu_count() or 1
        cv, cms, ckind = reference_train_rate(st, K, mean, std, "cpu", 120, 5, threads=cores)
        out["cpu_baseline"] = {"value": cv, "unit": "events/s", "cores": cores, "kind": ckind,
                               "sample": "120 consecutive 200-event training batches of the same stream after 5 warm-up "
                                         "batches (the reference's own train_link.train)" if ckind == "reference" else
                                         "120 consecutive 200-event training batches of the same stream (oracle port)"}
        # single-GPU point of the TGB-eval scaling curve (scaling runs report this metric as `value`)
        import types
        from bench_tgb import run_tgb_eval_bench
        torch.cuda.empty_cache()
        r = run_tgb_eval_bench(types.SimpleNamespace(steps=max(args.steps, 100), warmup=W, gpus=1), 0, 1, quiet=True,
                               extras=False)
        out["tgb_eval"] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "mrr", "rank0_subgraph", "e2e",
                                             "launches_per_step", "steps")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
