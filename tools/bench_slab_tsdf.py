"""BASELINE.json configs[4], TSDF half: one volume partitioned into Z-slabs over the ranks of a node.
Every rank integrates the same scan stream into ITS slab (+1 halo block), ray-casts its slab, and the renders are combined by
ONE per-pixel nearest-hit MIN all-reduce over NCCL on device buffers (tandem_b200.parallel.reduce_nearest_hit_device).

    python tools/bench_slab_tsdf.py [--frames 12]                                                   # 1 GPU (no collective)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_slab_tsdf.py --frames 12

Rank 0 also fuses the same stream into a single un-partitioned volume and checks the combined render against it (same
surface; sample positions differ where a ray restarts in another slab).  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tandem_b200 import DrFusion, DrFusionOptions  # noqa: E402
from tandem_b200.parallel import attach_peers, reduce_max, reduce_nearest_hit_device, slab_bounds, stream_barrier  # noqa: E402
from tandem_b200.synthetic import RoomScene, circle_trajectory  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=2, help="untimed frames (same stream, integrated before the timed ones)")
    ap.add_argument("--interleave", type=int, default=0, help="interleaved slabs of this many block rows (0: contiguous slabs)")
    ap.add_argument("--peer", action="store_true", help="pixel-partitioned ray-cast over the slabs: P2P voxel reads inside the ray-cast kernel (bit-identical to one volume)")
    ap.add_argument("--legacy", action="store_true", help="round-1 path: unclipped slab ray-cast, per-slab D2H, pack kernel + host syncs")
    a = ap.parse_args(argv)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)          # NCCL's version banner goes to stderr, stdout carries the JSON line only
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    H, W = 480, 640
    intr = dict(fx=320.0, fy=320.0, cx=319.5, cy=239.5)
    half = 5.0                                   # 10 m room inside the 10.24 m (1024^3 voxels at 1 cm) cube
    off = np.float32(5.12)
    scene = RoomScene(half=half, spheres=((2.4, 0.6, 1.6, 1.0), (-2.0, -0.8, 3.0, 1.0), (0.4, 1.8, -2.8, 1.0)))
    poses = circle_trajectory(a.frames + a.warmup, radius=2.0)
    frames = [scene.render(p, H, W, **intr, noise_sigma=0.002, dropout=0.02, seed=k) for k, p in enumerate(poses)]
    for p in poses:
        p[:3, 3] += off
    zmin, zmax = 0, 128                          # voxel blocks (8 cm) covering z in [0, 10.24)
    lo, hi, alo, ahi = slab_bounds(zmin, zmax, rank, world)
    # the reference allocates every block between the camera and the surface (tsdf_volume.cu:317-434), i.e. the whole room:
    # 1000 m^3 / (8 cm)^3 = 1.95 M blocks -> 2.5 M blocks (10 GB of voxels) instead of initDr's 1 M
    opt = DrFusionOptions(height=H, width=W, num_blocks=2500000, num_buckets=2500000, **intr)
    f = DrFusion(opt, device=local)
    if world > 1 and a.peer:
        f.set_slab(-(1 << 19) if rank == 0 else lo, (1 << 19) if rank == world - 1 else hi)   # owned rows only, no halo
        attach_peers(dist, f, rank, world)
    elif world > 1:
        if a.interleave > 0 and not a.legacy:
            f.set_interleave(rank, world, a.interleave, zmin)
        else:
            f.set_slab(alo, ahi)
    if a.legacy:
        f.set_option("slab_clip", 0)
    else:
        f.set_option("slab_exchange", 1)      # ray-cast emits the exchange keys, no per-slab D2H, one host sync per frame
    dev = f"cuda:{local}"
    from tandem_b200._lib import pinned_empty
    n_ring = len(poses) if (rank == 0 and world > 1) else 2   # rank 0 keeps every combined render for the parity check below
    ring = [(pinned_empty((H, W), np.float32), pinned_empty((H, W, 3), np.uint8)) for _ in range(n_ring)]

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local)

    if not a.legacy:   # the caller keeps its scans in page-locked memory: IntegrateScanAsync DMA's straight from them
        pf = []
        for bgr, depth in frames:
            pb, pd = pinned_empty(bgr.shape, np.uint8), pinned_empty(depth.shape, np.float32)
            pb[...] = bgr; pd[...] = depth
            pf.append((pb, pd))
        frames_in = pf
    else:
        frames_in = frames
    outs = []
    for k, ((bgr, depth), pose) in enumerate(zip(frames_in, poses)):
        if k == a.warmup:
            sync()
            t0 = time.perf_counter()
        f.IntegrateScanAsync(bgr, depth, pose)
        if a.peer:
            stream_barrier(dist, f, dev)        # every rank's integration of this scan precedes every rank's peer reads
        f.RenderAsync([pose])
        f.GetRenderResult()
        outs.append(reduce_nearest_hit_device(dist, f, 0, dev, out=None if a.legacy else ring[k % n_ring]))
    sync()
    ms = (time.perf_counter() - t0) * 1e3 / a.frames
    mi, mr = f.run_resident(5)                       # device time of the last frame's integrate / ray-cast on this rank
    mi, mr = reduce_max(dist, [mi / 5, mr / 5], device=dev)
    (ms,) = reduce_max(dist, [ms], device=dev)
    st = f.stats()
    blocks = reduce_max(dist, [st["allocated_blocks"]], device=dev)[0]
    if rank == 0:
        line = {"what": "TSDF Z-slab partition, integrate + ray-cast + nearest-hit all-reduce (NCCL, device buffers)", "n_gpus": world,
                "frames": a.frames, "ms_per_frame(max over ranks, wall incl. H2D/D2H)": ms, "frames_per_s": 1e3 / ms,
                "max_blocks_per_rank": int(blocks), "slab_blocks(owned, rank 0)": [lo, hi],
                "device_ms_last_frame(max over ranks)": {"allocate+integrate": mi, "raycast": mr},
                "mode": "legacy (round 1)" if a.legacy else "slab-clipped ray-cast + fused key packing + all-reduce on the fusion stream",
                "partition": "contiguous Z-slabs, pixel-partitioned ray-cast with P2P voxel reads" if a.peer else "contiguous Z-slabs" if (a.legacy or a.interleave <= 0) else f"interleaved Z-slabs of {a.interleave} block rows (+1 halo row each side)"}
        if world > 1:     # parity against the un-partitioned volume
            full = DrFusion(opt, device=local)
            mism, med, p99 = [], [], []
            for k, ((bgr, depth), pose) in enumerate(zip(frames, poses)):
                full.IntegrateScanAsync(bgr, depth, pose)
                full.RenderAsync([pose])
                (fb,), (fd,) = full.GetRenderResult()
                dm, bm = outs[k]
                hm, hf = dm > 0, fd > 0
                both = hm & hf
                mism.append(float(np.mean(hm != hf)))
                err = np.abs(dm[both] - fd[both])
                med.append(float(np.median(err))); p99.append(float(np.quantile(err, 0.99)))
            line.update({"single_volume_blocks": full.stats()["allocated_blocks"], "hit_mismatch_max": max(mism),
                         "depth_abs_err_median_max_m": max(med), "depth_abs_err_p99_max_m": max(p99)})
            assert max(mism) < 5e-3 and max(med) < 1e-3 and max(p99) < 0.03, line
            if a.peer:
                assert max(mism) == 0 and max(p99) == 0, ("the peer ray-cast must reproduce the single volume exactly", line)
        print(json.dumps(line))
        result = line
    else:
        result = None
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return result


if __name__ == "__main__":
    main()
