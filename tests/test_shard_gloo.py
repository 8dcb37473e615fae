"""CPU, world_size 2 over gloo: the multi-GPU path of bench.py is pure stream sharding (no
collective on the data path) plus a max-over-ranks reduction of the step time."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out):
    sys.path.insert(0, ROOT)
    import bench

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bench.shard(n_total, world, rank)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(mine)]))
    pad = torch.full((max(int(s) for s in sizes),), -1, dtype=torch.int64)
    pad[: len(mine)] = torch.tensor(mine, dtype=torch.int64)
    allv = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(allv, pad)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)  # per-rank step time -> max over ranks
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        got = sorted(int(x) for v in allv for x in v if x >= 0)
        out.put((got, float(t)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [512, 7])
def test_stream_sharding_two_ranks(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == list(range(n_total))  # every stream owned by exactly one rank
    assert tmax == 11.0


def test_shard_is_contiguous_and_balanced():
    sys.path.insert(0, ROOT)
    import bench

    for n, w in ((256, 1), (2048, 8), (10, 4), (3, 8)):
        parts = [bench.shard(n, w, r) for r in range(w)]
        assert sum(parts, []) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
