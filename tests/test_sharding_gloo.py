"""CPU, world_size 2 over gloo: the multi-GPU path is pure sharding (env i -> rank i // per) plus one
SUM all-reduce of the statistics vector and a MAX all-reduce of the device time."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphenvs_b200.sharding import max_over_ranks, rank_slice, reduce_stats


def test_rank_slice_partitions_exactly():
    for total in (1, 7, 64, 65536, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [rank_slice(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for lo, c in spans:
                if c:
                    assert lo == pos
                pos += c


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, cnt = rank_slice(1000, rank, world)
    per_env = torch.arange(lo, lo + cnt, dtype=torch.float64)
    stats = torch.stack([torch.tensor(float(cnt), dtype=torch.float64), (per_env % 3 == 0).sum().double(), per_env.sum(),
                         (per_env * 0.5).sum()])
    reduce_stats(stats)
    tmax = max_over_ranks(10.0 + rank)
    q.put((rank, stats.tolist(), tmax))
    dist.destroy_process_group()


def test_stats_allreduce_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = torch.arange(1000, dtype=torch.float64)
    exp = [1000.0, float((full % 3 == 0).sum()), float(full.sum()), float((full * 0.5).sum())]
    for _, stats, tmax in res:
        assert stats == pytest.approx(exp)
        assert tmax == 11.0
