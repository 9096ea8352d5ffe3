"""world_size-2 CPU test (gloo) of the N>1 host logic bench.py uses: contiguous frame sharding with no data-path
collective, barrier, and max-over-ranks timing reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nframes, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import framewright_b200  # noqa: F401
    from framewright_b200.multi_gpu import shard_range

    lo, hi = shard_range(nframes, world, rank)
    # each rank "processes" its own frames only: per-frame checksum of a frame generated from (seed, index)
    mine = torch.zeros(nframes, dtype=torch.int64)
    for i in range(lo, hi):
        g = torch.Generator().manual_seed(1234 + i)
        mine[i] = torch.randint(0, 256, (16,), generator=g).sum()
    dist.barrier()
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)          # pretend device time
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                       # bench.py: max over ranks
    counts = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(counts)                                         # test-only: verify the partition is exact
    dist.all_reduce(mine)
    if rank == 0:
        out_q.put((float(ms.item()), int(counts.item()), mine.tolist()))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduce():
    world, nframes = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    ms, count, sums = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 11.0 and count == nframes
    want = []
    for i in range(nframes):
        g = torch.Generator().manual_seed(1234 + i)
        want.append(int(torch.randint(0, 256, (16,), generator=g).sum()))
    assert sums == want
