"""Device NegativeSampler (csrc/sampler.cu) against the reference's sampler (negative_sampler.py:8-48, imported
unmodified when available): same support, no collision with the seen-edge set, uniform over the unseen destinations,
eval-mode determinism.  The numpy RNG stream is not reproduced, so draws are compared as distributions."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _edges(seed=0, n=4000, n_src=40, n_dst=25):
    rng = np.random.default_rng(seed)
    src = torch.from_numpy(rng.integers(0, n_src, n))
    dst = torch.from_numpy(rng.integers(100, 100 + n_dst, n))
    dst[dst == 107] = 108                                    # a gap in the id range: 107 is not a destination node
    return src, dst


def test_no_collisions_support_and_uniformity():
    from ctan_b200.negative_sampler import NegativeSampler
    src, dst = _edges(n=300)                                 # sparse: every source has unseen destinations
    smp = NegativeSampler(src, dst, name="t", seed=9, device=DEV)
    seen = set(zip(src.tolist(), dst.tolist()))
    dst_set = set(dst.tolist())
    q = torch.full((20000,), 3, dtype=torch.long, device=DEV)
    draws = smp.sample(q)
    assert draws.device == q.device and draws.dtype == torch.long and draws.shape == q.shape
    vals = draws.cpu().tolist()
    assert set(vals) <= dst_set and 107 not in vals
    assert all((3, v) not in seen for v in vals)
    allowed = sorted(d for d in dst_set if (3, d) not in seen)
    counts = np.array([vals.count(d) for d in allowed], dtype=np.float64)
    expect = len(vals) / len(allowed)
    assert np.all(np.abs(counts - expect) < 6 * np.sqrt(expect)), (counts, expect)     # uniform over the unseen ones
    assert smp.failed_draws() == 0
    # successive training calls differ; eval calls repeat (fresh generator per call upstream)
    a, b = smp.sample(q[:500]), smp.sample(q[:500])
    assert not torch.equal(a, b)
    e1, e2 = smp.sample(q[:500], eval=True, eval_seed=12345), smp.sample(q[:500], eval=True, eval_seed=12345)
    assert torch.equal(e1, e2) and not torch.equal(e1, smp.sample(q[:500], eval=True, eval_seed=7))


def test_attempt_cap_and_no_check_mode():
    from ctan_b200.negative_sampler import NegativeSampler
    src = torch.zeros(30, dtype=torch.long)
    dst = torch.arange(30) + 5                               # source 0 has seen EVERY destination
    smp = NegativeSampler(src, dst, name="t", device=DEV)
    out = smp.sample(torch.zeros(64, dtype=torch.long, device=DEV))
    assert smp.failed_draws() == 64 and set(out.cpu().tolist()) <= set(dst.tolist())   # capped, as upstream
    free = NegativeSampler(src, dst, name="t", check_link_existence=False, device=DEV)
    assert free.seen is None and free.sample(torch.zeros(64, dtype=torch.long, device=DEV)).numel() == 64


def test_same_distribution_as_the_reference_sampler():
    from oracle import ref_loops
    if not ref_loops.available():
        pytest.skip("reference files not present")
    R = ref_loops.load()
    from ctan_b200.negative_sampler import NegativeSampler
    src, dst = _edges(seed=2, n=600)
    ref = R.negative_sampler.NegativeSampler(src, dst, name="r", seed=9)
    mine = NegativeSampler(src, dst, name="m", seed=9, device=DEV)
    assert torch.equal(mine.dst_nodes.cpu(), ref.dst_nodes) and torch.equal(mine.src_nodes.cpu(), ref.src_nodes)
    q = torch.from_numpy(np.random.default_rng(1).integers(0, 40, 3000))
    a = ref.sample(q)
    b = mine.sample(q.to(DEV)).cpu()
    seen = set(zip(src.tolist(), dst.tolist()))
    for s, x, y in zip(q.tolist(), a.tolist(), b.tolist()):
        assert (s, x) not in seen and (s, y) not in seen
    ha = np.bincount(a.numpy() - 100, minlength=25).astype(np.float64)
    hb = np.bincount(b.numpy() - 100, minlength=25).astype(np.float64)
    assert np.all((ha == 0) == (hb == 0))
    nz = ha > 0
    assert np.all(np.abs(ha[nz] - hb[nz]) < 6 * np.sqrt(ha[nz] + hb[nz]))            # same histogram up to sampling noise
