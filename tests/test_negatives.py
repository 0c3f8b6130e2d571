"""f2: CSR-packed negative supply against the reference's NegativeEdgeSampler.query_batch semantics
(TGB/tgb/linkproppred/negative_sampler.py:83-136, restated here: a dict lookup per positive edge)."""
import numpy as np
import pytest
import torch

import ctan_b200  # noqa: F401
from ctan_b200.negatives import PackedNegatives


def _query_batch_ref(eval_set, src, dst, t):
    out = []
    for s, d, tt in zip(src.tolist(), dst.tolist(), t.tolist()):
        if (s, d, tt) not in eval_set:
            raise ValueError("missing")
        out.append([int(x) for x in eval_set[(s, d, tt)]])
    return out


def _make(n=500, ragged=False, seed=0):
    rng = np.random.default_rng(seed)
    src = torch.from_numpy(rng.integers(0, 50, n)); dst = torch.from_numpy(rng.integers(50, 90, n))
    t = torch.from_numpy(np.sort(rng.integers(0, 10_000, n)))
    ev = {}
    for s, d, tt in zip(src.tolist(), dst.tolist(), t.tolist()):
        k = int(rng.integers(1, 9)) if ragged else 7
        ev.setdefault((s, d, tt), rng.integers(50, 90, k))
    return ev, src, dst, t


def test_packed_matches_query_batch():
    ev, src, dst, t = _make()
    pk = PackedNegatives.from_eval_set(ev, src, dst, t)
    assert pk.is_uniform() and pk.dense().shape == (500, 7)
    for lo in range(0, 500, 200):
        hi = min(lo + 200, 500)
        assert pk.query_batch(lo, hi) == _query_batch_ref(ev, src[lo:hi], dst[lo:hi], t[lo:hi])
        rp, neg = pk.rows(lo, hi)
        assert rp[0] == 0 and rp[-1] == neg.numel() == 7 * (hi - lo)
        assert torch.equal(neg.view(-1, 7), pk.dense()[lo:hi])


def test_ragged_and_errors():
    ev, src, dst, t = _make(ragged=True, seed=3)
    pk = PackedNegatives.from_eval_set(ev, src, dst, t)
    assert pk.query_batch(0, 500) == _query_batch_ref(ev, src, dst, t)
    if not pk.is_uniform():
        with pytest.raises(ValueError):
            pk.dense()
    bad = dict(ev)
    bad.pop(next(iter(bad)))
    with pytest.raises(ValueError):
        PackedNegatives.from_eval_set(bad, src, dst, t)
    with pytest.raises(ValueError):
        PackedNegatives.from_eval_set(None, src, dst, t)
