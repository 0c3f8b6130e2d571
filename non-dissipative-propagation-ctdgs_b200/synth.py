"""Synthetic continuous-time event streams shaped like the reference's datasets
(SURVEY.md S8(d)); there is no network for the real JODIE / TGB files.

Shapes follow the reference's loaders: JODIE Wikipedia/Reddit (`datasets/__init__.py:36-43`:
bipartite, 172-d messages, `x = rng.random((num_nodes, 1))`), TGB (`train_link.py:458`: `x = ones`),
tgbl-comment (`TGB/tgb/utils/pre_process.py:438`: non-bipartite reply network, 2-d messages)."""
from __future__ import annotations

import numpy as np
import torch

SHAPES = {
    # name: (n_users, n_items, n_events, msg_dim, bipartite)
    "wiki": (8227, 1000, 157474, 172, True),
    "reddit": (10000, 984, 672447, 172, True),
    "comment": (994790, 0, 44314507, 2, False),
}
T_MAX = 2678373


def _zipf_ids(rng, n, count, a):
    ranks = np.arange(1, n + 1, dtype=np.float64) ** (-a)
    p = ranks / ranks.sum()
    ids = rng.choice(n, size=count, p=p)
    perm = rng.permutation(n)
    return perm[ids].astype(np.int64)


def make_stream(shape: str = "wiki", n_events: int | None = None, seed: int = 0, zipf_a: float = 1.2,
                tgb: bool = False, msg_dim: int | None = None, n_nodes_scale: float = 1.0) -> dict:
    """Returns a dict of CPU tensors: src,dst,t (int64), msg (fp32 [n,F_e]), x (fp32 [num_nodes,1]),
    plus num_nodes, min_dst, max_dst."""
    n_users, n_items, n_full, f_e, bip = SHAPES[shape]
    n_users = max(2, int(n_users * n_nodes_scale))
    n_items = int(n_items * n_nodes_scale) if bip else 0
    n = n_full if n_events is None else n_events
    f_e = msg_dim or f_e
    rng = np.random.default_rng(seed)
    src = _zipf_ids(rng, n_users, n, zipf_a)
    if bip:
        dst = _zipf_ids(rng, n_items, n, zipf_a) + n_users
        num_nodes = n_users + n_items
    else:
        dst = _zipf_ids(rng, n_users, n, zipf_a)
        num_nodes = n_users
    t = np.sort(rng.integers(0, T_MAX + 1, size=n)).astype(np.int64)
    g = torch.Generator().manual_seed(seed)
    msg = torch.randn(n, f_e, generator=g)
    x = torch.ones(num_nodes, 1) if tgb else torch.from_numpy(rng.random((num_nodes, 1)).astype(np.float32))
    return {"src": torch.from_numpy(src), "dst": torch.from_numpy(dst), "t": torch.from_numpy(t), "msg": msg,
            "x": x, "num_nodes": num_nodes, "min_dst": int(dst.min()), "max_dst": int(dst.max())}


def draw_negatives(stream: dict, per_event: int = 1, seed: int = 1) -> torch.Tensor:
    """Negatives drawn once on the host, uniform over the destination id range
    (`train_link.py:245-251`), shape [n_events, per_event]; fed to every implementation."""
    g = torch.Generator().manual_seed(seed)
    n = stream["src"].numel()
    return torch.randint(stream["min_dst"], stream["max_dst"] + 1, (n, per_event), generator=g, dtype=torch.long)


def delta_t_stats(stream: dict, n_train: int, init_time: int = 0):
    """mean/std of inter-event times per node over the first n_train events, with the semantics of
    `utils.py:46-92` (both endpoints contribute; first occurrence measured from init_time)."""
    src = stream["src"][:n_train].numpy(); dst = stream["dst"][:n_train].numpy(); t = stream["t"][:n_train].numpy()
    last = {}
    diffs = np.empty(2 * n_train, dtype=np.float64)
    for k in range(n_train):
        s, d, tt = int(src[k]), int(dst[k]), int(t[k])
        diffs[2 * k] = tt - last.get(s, init_time)
        diffs[2 * k + 1] = tt - last.get(d, init_time)
        last[s] = tt
        last[d] = tt
    return float(diffs.mean()), float(diffs.std())
