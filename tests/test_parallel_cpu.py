"""world_size-2 gloo test of the N>1 plumbing used by bench.py (window sharding + max/sum reductions)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tandem_b200.parallel import aggregate_throughput, reduce_max, windows_for_rank


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = windows_for_rank(7, rank, world)
    ms = 10.0 * (rank + 1)                       # rank 1 is the slow one
    thr = aggregate_throughput(dist, len(mine), ms)
    mx = reduce_max(dist, [ms, float(rank)])
    dist.barrier()
    q.put((rank, mine, thr, mx))
    dist.destroy_process_group()


def test_window_sharding_covers_everything_once():
    for world in (1, 2, 3, 8):
        seen = sorted(w for r in range(world) for w in windows_for_rank(13, r, world))
        assert seen == list(range(13))
    with pytest.raises(ValueError):
        windows_for_rank(4, 2, 2)


def test_two_rank_gloo_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    for r in res:
        assert abs(r[2] - 7 / 0.020) < 1e-6          # 7 windows / max(10, 20) ms
        assert r[3] == [20.0, 1.0]
