"""Host-side logic of the sharded TGB evaluation on CPU: query partition + the single all-gather
exchange (world_size 2 and 3, gloo)."""
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_every_batch():
    sys.path.insert(0, ROOT)
    import ctan_b200  # noqa: F401
    from ctan_b200.tgb_eval import shard_range
    for B in (1, 7, 200, 201):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                q0, q1 = shard_range(B, r, world)
                assert 0 <= q0 <= q1 <= B and (q1 - q0) <= shard_range(B, 0, world)[1]
                cover += list(range(q0, q1))
            assert cover == list(range(B))


def _worker(rank, world, port, B, row, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import ctan_b200  # noqa: F401
    from ctan_b200.tgb_eval import exchange_rows, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    q0, q1 = shard_range(B, rank, world)
    qmax = shard_range(B, 0, world)[1]
    send = torch.zeros(qmax, row)
    for q in range(q0, q1):                       # row content encodes the global query id
        send[q - q0] = torch.arange(row, dtype=torch.float32) + 1000.0 * q
    recv = torch.zeros(world * qmax, row)
    rows = torch.zeros(B, row)
    exchange_rows(send, recv, rows, B, world)
    expect = torch.stack([torch.arange(row, dtype=torch.float32) + 1000.0 * q for q in range(B)])
    ok = torch.equal(rows, expect)
    gathered = [None] * world
    dist.all_gather_object(gathered, bool(ok))
    if rank == 0:
        out.put(all(gathered))
    dist.destroy_process_group()


def _run(world, B):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + world * 7 + B % 50
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, 11, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=10) is True


def test_exchange_rows_world2():
    _run(2, 200)


def test_exchange_rows_world3_ragged():
    _run(3, 7)
