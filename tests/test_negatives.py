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


def test_packed_matches_the_reference_sampler_itself():
    """Against the reference's own NegativeEdgeSampler.query_batch (TGB/tgb/linkproppred/negative_sampler.py:83-138,
    imported unmodified): same lists for every batch, same ValueError for an edge missing from the set."""
    from oracle import ref_loops
    if not ref_loops.available():
        pytest.skip("reference files not present (neither /root/reference nor baseline/_ref/)")
    R = ref_loops.load()
    for ragged in (False, True):
        ev, src, dst, t = _make(ragged=ragged, seed=11)
        smp = R.NegativeEdgeSampler(dataset_name="tgbl-wiki", strategy="hist_rnd")
        smp.eval_set["val"] = ev
        pk = PackedNegatives.from_eval_set(ev, src, dst, t)
        for lo in range(0, 500, 200):
            hi = min(lo + 200, 500)
            assert pk.query_batch(lo, hi) == smp.query_batch(src[lo:hi], dst[lo:hi], t[lo:hi], split_mode="val")
        if not ragged:
            want = torch.tensor(smp.query_batch(src, dst, t, split_mode="val"))
            assert torch.equal(pk.dense(), want)
    bad = dict(ev)
    k = (int(src[0]), int(dst[0]), int(t[0]))
    bad.pop(k)
    smp.eval_set["val"] = bad
    with pytest.raises(ValueError):
        smp.query_batch(src[:1], dst[:1], t[:1], split_mode="val")
    with pytest.raises(ValueError):
        PackedNegatives.from_eval_set(bad, src, dst, t)
