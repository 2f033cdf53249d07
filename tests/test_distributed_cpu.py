"""world_size-2 gloo test (CPU) of the sharding / gather logic used on the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mapdn_b200.distributed import all_gather_varlen, shard_range


def test_shard_range_partitions():
    for gb in (1, 7, 8, 1024, 8191):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(gb, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == gb
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, world, port, gb, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, cnt = shard_range(gb, rank, world)
    local = torch.arange(off, off + cnt, dtype=torch.float64) * 1.5     # "episode return" of global env id
    g = all_gather_varlen(local, gb)
    info = torch.ones(cnt, 11, dtype=torch.float64) * (rank + 1)
    tot = info.sum(0)
    dist.all_reduce(tot)
    if rank == 0:
        q.put((g.tolist(), tot.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("gb", [8, 7])
def test_gather_world2_gloo(gb):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gb, q)) for r in range(2)]
    for p in procs:
        p.start()
    g, tot = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert g == [1.5 * i for i in range(gb)]
    c0, c1 = shard_range(gb, 0, 2)[1], shard_range(gb, 1, 2)[1]
    assert tot == [c0 * 1.0 + c1 * 2.0] * 11
