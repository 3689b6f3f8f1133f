"""CPU, world_size 2 over gloo: the N>1 path (sample sharding like the reference's get_chunk, token-id gather,
max-over-ranks timing). The hot path itself has no collective (replicas only)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llava._b2.replicas import gather_rows, max_over_ranks, shard_range, sum_over_ranks


def test_shard_range_covers_everything():
    for n in (0, 1, 5, 8, 31, 32, 33):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
            assert got == list(range(n))


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    full = torch.arange(total * 3, dtype=torch.int32).reshape(total, 3)
    mine = full[lo:hi].clone()
    gathered = gather_rows(mine, total)
    ok = bool((gathered == full).all())
    t = max_over_ranks(1.0 + rank)
    s = sum_over_ranks(hi - lo)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, t, s))


def test_two_rank_gather_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    for total in (5, 8):  # ragged and even shards
        procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        for rank, ok, t, s in res:
            assert ok and t == 2.0 and s == total
        port += 1
