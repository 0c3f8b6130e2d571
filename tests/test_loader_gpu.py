"""Bit-exact parity of the ring-buffer kernels (csrc/loader.cu) with the oracle's restatement of
LastNeighborLoader (TGB/modules/neighbor_loader.py:15-85) and SimpleMemory.update_state
(models/memory_layers.py:132-144).  Calls go through the C-ABI library."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cmp_tables(mine, ref):
    e_m, e_r = mine.e_id.cpu(), ref.e_id
    assert torch.equal(e_m, e_r)
    valid = e_r >= 0
    assert torch.equal(mine.neighbors.cpu()[valid], ref.neighbors[valid])
    assert torch.equal(mine._deg.cpu().long(), valid.sum(1))


def _replay(src, dst, num_nodes, S, B, hops=(1,), neg=None, check_every=1):
    from ctan_b200.models import LastNeighborLoader
    from oracle import ctan_oracle as O
    dev = torch.device("cuda:0")
    mine = LastNeighborLoader(num_nodes, S, device=dev)
    ref = O.TorchNeighborLoader(num_nodes, S)
    n = src.numel()
    for it, lo in enumerate(range(0, n, B)):
        s, d = src[lo:lo + B], dst[lo:lo + B]
        seeds = torch.cat([s, d] + ([neg[lo:lo + B]] if neg is not None else []))
        if it % check_every == 0:
            for K in hops:
                a = O.khop_lookup(ref, seeds.unique(), K)
                n_id, ei, eid, rowptr = mine.lookup_khop(seeds.to(dev), K)     # duplicates allowed
                assert torch.equal(n_id.cpu(), a[0]), f"n_id mismatch batch {it} K={K}"
                assert torch.equal(ei.cpu(), a[1]), f"edge_index mismatch batch {it} K={K}"
                assert torch.equal(eid.cpu(), a[2]), f"e_id mismatch batch {it} K={K}"
                cnt = torch.bincount(a[1][1], minlength=a[0].numel()) if a[1].numel() else torch.zeros(a[0].numel(), dtype=torch.long)
                assert torch.equal(rowptr.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), cnt.cumsum(0)]))
            # the reference call chain itself (one hop per call)
            b = mine(seeds.unique().to(dev))
            a = ref(seeds.unique())
            for x, y in zip(b, a):
                assert torch.equal(x.cpu(), y)
        mine.insert(s.to(dev), d.to(dev))
        ref.insert(s, d)
        assert mine.cur_e_id == ref.cur_e_id
        if it % check_every == 0:
            _cmp_tables(mine, ref)
    _cmp_tables(mine, ref)
    return mine, ref


@pytest.mark.parametrize("S,hops", [(5, (1, 2, 3)), (32, (1, 2)), (3, (1, 5))])
def test_replay_wiki_shaped(small_stream, S, hops):
    st = small_stream
    _replay(st["src"], st["dst"], st["num_nodes"], S, 200, hops=hops, neg=st["neg"])


def test_overflow_hub_and_self_loops():
    """One hub with far more than S events per batch (as src AND as dst), plus self-loops."""
    g = torch.Generator().manual_seed(3)
    n, num_nodes = 1500, 64
    src = torch.randint(0, num_nodes, (n,), generator=g)
    dst = torch.randint(0, num_nodes, (n,), generator=g)
    hub = torch.rand(n, generator=g)
    src[hub < 0.3] = 7
    dst[(hub > 0.25) & (hub < 0.5)] = 7
    loops = torch.rand(n, generator=g) < 0.05
    dst[loops] = src[loops]
    for S in (1, 2, 5, 8):
        _replay(src, dst, num_nodes, S, 100, hops=(1, 2))


@pytest.mark.parametrize("B", [1, 7, 33, 600])
def test_batch_sizes_and_ragged_tail(B):
    g = torch.Generator().manual_seed(5)
    n, num_nodes = 1237, 300
    src = torch.randint(0, 200, (n,), generator=g)
    dst = torch.randint(200, num_nodes, (n,), generator=g)
    _replay(src, dst, num_nodes, 4, B, hops=(1, 2), check_every=1 if B > 1 else 50)


def test_large_sampler_size():
    g = torch.Generator().manual_seed(6)
    n, num_nodes = 3000, 50
    src = torch.randint(0, num_nodes, (n,), generator=g)
    dst = torch.randint(0, num_nodes, (n,), generator=g)
    _replay(src, dst, num_nodes, 40, 250, hops=(1,))


def test_empty_state_and_empty_inputs():
    from ctan_b200.models import LastNeighborLoader
    dev = torch.device("cuda:0")
    ld = LastNeighborLoader(100, 5, device=dev)
    n_id, ei, eid = ld(torch.tensor([3, 9, 50], device=dev))
    assert n_id.tolist() == [3, 9, 50] and ei.shape == (2, 0) and ei.dtype == torch.long and eid.numel() == 0
    n_id, ei, eid = ld(torch.empty(0, dtype=torch.long, device=dev))
    assert n_id.numel() == 0 and ei.shape == (2, 0)
    ld.insert(torch.empty(0, dtype=torch.long, device=dev), torch.empty(0, dtype=torch.long, device=dev))
    assert ld.cur_e_id == 0
    ld.insert(torch.tensor([1], device=dev), torch.tensor([2], device=dev))
    ld.reset_state()
    assert ld.cur_e_id == 0 and bool((ld.e_id == -1).all())
    n_id, ei, eid = ld(torch.tensor([1, 2], device=dev))
    assert ei.shape == (2, 0)


def test_million_node_table():
    """tgbl-comment-sized id space (994k nodes): bitmap scan over 31k words."""
    g = torch.Generator().manual_seed(7)
    num_nodes, n = 994790, 4000
    src = torch.randint(0, num_nodes, (n,), generator=g)
    dst = torch.randint(0, num_nodes, (n,), generator=g)
    src[::3] = src[0]
    _replay(src, dst, num_nodes, 8, 500, hops=(1, 2), check_every=2)


def test_memory_update_matches_oracle():
    from ctan_b200.models import SimpleMemory
    from oracle import ctan_oracle as O
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    num_nodes, D = 300, 24
    mine = SimpleMemory(num_nodes, D, None, init_time=0).to(dev)
    ref_mem, ref_lu = torch.zeros(num_nodes, D), torch.zeros(num_nodes, dtype=torch.long)
    for it in range(12):
        B = [1, 5, 64, 257][it % 4]
        src = torch.randint(0, 40, (B,), generator=g)
        dst = torch.randint(0, 40, (B,), generator=g)
        t = torch.randint(0, 6, (B,), generator=g) + 10 * it      # many ties inside a batch
        if it % 3 == 0:
            t = torch.sort(t).values
        se, de = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
        O.memory_update_ref(ref_mem, ref_lu, src, dst, t, se, de)
        mine.update_state(src.to(dev), dst.to(dev), t.to(dev), se.to(dev), de.to(dev))
        assert torch.equal(mine.last_update.cpu(), ref_lu)
        assert torch.equal(mine.memory.cpu(), ref_mem)          # rows are copied: bit-exact
    mine.reset_state()
    assert float(mine.memory.abs().sum()) == 0 and int(mine.last_update.abs().sum()) == 0
