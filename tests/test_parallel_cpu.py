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


def _slab_worker(rank, world, port, q):
    import numpy as np
    from tandem_b200.parallel import reduce_nearest_hit
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                         # same "scene" on both ranks
    full = rng.uniform(0.5, 4.0, (12, 16)).astype(np.float32)
    col = rng.integers(0, 255, (12, 16, 3)).astype(np.uint8)
    owner = rng.integers(0, 2, (12, 16))                   # which slab holds the nearest surface of each ray
    miss = rng.random((12, 16)) < 0.2
    mine = np.where((owner == rank) & ~miss, full, 0).astype(np.float32)   # this rank only sees hits inside its slab
    far = rng.random((12, 16)) < 0.3                       # ... plus some farther hits behind another rank's surface
    mine = np.where((owner != rank) & far & ~miss, full + 1.0, mine).astype(np.float32)
    d, c = reduce_nearest_hit(dist, mine, col)
    q.put((rank, d, c, full, col, miss))
    dist.destroy_process_group()


def test_slab_render_reduction_gloo():
    import numpy as np
    from tandem_b200.parallel import pack_hits, slab_bounds, unpack_hits
    assert slab_bounds(-32, 32, 0, 2) == (-32, 0, -33, 1) and slab_bounds(-32, 32, 1, 2) == (0, 32, -1, 33)
    covered = sorted(z for r in range(3) for z in range(*slab_bounds(-5, 6, r, 3)[:2]))
    assert covered == list(range(-5, 6))
    d = np.array([[0.0, 1.5], [2.25, 0.0]], np.float32)
    c = np.array([[[0, 0, 0], [1, 2, 3]], [[250, 251, 252], [9, 9, 9]]], np.uint8)
    d2, c2 = unpack_hits(pack_hits(d, c))
    assert np.array_equal(d, d2) and np.array_equal(c2[0, 1], [1, 2, 3]) and np.array_equal(c2[1, 1], [0, 0, 0])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_slab_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, dd, cc, full, col, miss in res:
        assert np.array_equal(dd, np.where(miss, 0, full).astype(np.float32))
        assert np.array_equal(cc[~miss], col[~miss]) and not cc[miss].any()


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) works without a GPU and prints exactly one JSON
    line with the contract's keys; under torchrun only rank 0 prints."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]
    env["RANK"] = "1"
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=60, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
