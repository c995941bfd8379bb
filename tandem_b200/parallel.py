"""Multi-GPU plumbing for the paths that shard: independent keyframe windows over ranks.
No data-path collective exists on this path; torch.distributed is used only for barriers and for reducing the
per-rank timings / counters (NCCL on GPUs, gloo in the CPU tests)."""
from typing import List, Sequence


def windows_for_rank(n_windows: int, rank: int, world: int) -> List[int]:
    """Window w is owned by rank w mod world (round robin keeps the load balanced for any n_windows)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_windows, world))


def reduce_max(dist, values: Sequence[float], device="cpu") -> List[float]:
    """Element-wise max over ranks of a short list of floats (timings are reported as the slowest rank's)."""
    if dist is None:
        return [float(v) for v in values]
    import torch
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def reduce_sum(dist, values: Sequence[float], device="cpu") -> List[float]:
    if dist is None:
        return [float(v) for v in values]
    import torch
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def aggregate_throughput(dist, units_done: int, elapsed_ms: float, device="cpu") -> float:
    """Whole-job throughput = (units all ranks processed) / (max over ranks of the elapsed time)."""
    (total,) = reduce_sum(dist, [units_done], device)
    (ms,) = reduce_max(dist, [elapsed_ms], device)
    return total / (ms / 1e3)


# ---- TSDF Z-slab partition --------------------------------------------------------------------------

def slab_bounds(z_block_min: int, z_block_max: int, rank: int, world: int, halo: int = 1):
    """Blocks z in [z_block_min, z_block_max) are split into `world` contiguous slabs.  Returns (owned_lo, owned_hi,
    alloc_lo, alloc_hi): the rank OWNS [owned_lo, owned_hi) and additionally integrates `halo` blocks on either side so
    that trilinear reads near the slab faces see the same voxels as a single-volume run."""
    n = z_block_max - z_block_min
    lo = z_block_min + (n * rank) // world
    hi = z_block_min + (n * (rank + 1)) // world
    return lo, hi, lo - halo, hi + halo


def pack_hits(depth, bgr):
    """(depth f32 (H,W), bgr u8 (H,W,3)) -> int64 keys ordered by depth; a miss (depth == 0) becomes +inf.
    Positive IEEE floats order like their bit patterns, so min over keys == nearest hit, colour rides along."""
    import numpy as np
    d = np.ascontiguousarray(depth, np.float32)
    bits = d.view(np.uint32).astype(np.int64)
    bits = np.where(d > 0, bits, np.int64(0x7F800000))
    col = (bgr[..., 0].astype(np.int64) | (bgr[..., 1].astype(np.int64) << 8) | (bgr[..., 2].astype(np.int64) << 16))
    return (bits << 24) | col


def unpack_hits(keys):
    import numpy as np
    bits = (keys >> 24).astype(np.uint32)
    depth = bits.view(np.float32).copy()
    depth[bits == 0x7F800000] = 0.0
    col = keys & 0xFFFFFF
    bgr = np.stack([(col & 0xFF), (col >> 8) & 0xFF, (col >> 16) & 0xFF], -1).astype(np.uint8)
    bgr[depth == 0] = 0
    return depth, bgr


def reduce_nearest_hit(dist, depth, bgr, device="cpu"):
    """The one exchange step of the slab-partitioned ray-cast: per-pixel nearest hit over ranks (all-reduce MIN of packed
    keys, 2.46 MB at 640x480).  Valid because z-depth is monotone along a ray and the slabs are disjoint."""
    import torch
    keys = pack_hits(depth, bgr)
    if dist is None:
        return unpack_hits(keys)
    t = torch.from_numpy(keys).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return unpack_hits(t.cpu().numpy())


class _DeviceInt64:
    """__cuda_array_interface__ view of a device buffer owned by the C library (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def reduce_nearest_hit_device(dist, fusion, render_index=0, device="cuda:0", out=None):
    """Same exchange as reduce_nearest_hit but without the host round trip: the keys sit in a device buffer of the DrFusion
    handle (written by the ray-cast itself in "slab_exchange" mode, else by a pack kernel), are all-reduced in place with MIN
    over the ranks (NCCL over NVLink) and unpacked by a kernel.  The collective is enqueued ON THE HANDLE'S OWN STREAM
    (torch.cuda.ExternalStream), i.e. behind the ray-cast and in front of the unpack kernel, with no host synchronisation in
    between; the only sync of the frame is the D2H of the combined render inside unpack_keys.
    Returns (depth, bgr) of the combined render on the host (written into `out` = (depth, bgr) arrays when given - pass
    page-locked arrays for a copy-free read-back)."""
    ptr, n = fusion.render_keys_device(render_index)
    if dist is not None:
        import torch
        t = torch.as_tensor(_DeviceInt64(ptr, n), device=device)
        with torch.cuda.stream(torch.cuda.ExternalStream(fusion.stream(), device=device)):
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return fusion.unpack_keys(ptr, out)


def attach_peers(dist, fusion, rank, world):
    """Pixel-partitioned ray-cast over the Z-slab-partitioned volume: gather every rank's exported tables (CUDA IPC handles) and
    map them into this rank (NVLink P2P).  Call once, after set_slab(owned_lo, owned_hi) and before the first scan."""
    mine = fusion.peer_export()
    if dist is None or world == 1:
        fusion.peer_attach([mine], 0)
        return
    blobs = [None] * world
    dist.all_gather_object(blobs, mine)
    fusion.peer_attach(blobs, rank)


_barrier_token = {}


def stream_barrier(dist, fusion, device="cuda:0"):
    """Stream-ordered barrier over the ranks, enqueued on the DrFusion handle's own stream (a 1-element NCCL all-reduce): work
    enqueued on that stream afterwards - the ray-cast that reads the peers' voxels - runs only after every rank's earlier work
    on ITS stream - the integration of the same scan - has finished.  No host synchronisation."""
    if dist is None:
        return
    import torch
    t = _barrier_token.get(device)
    if t is None:
        t = _barrier_token[device] = torch.zeros(1, dtype=torch.int32, device=device)
    with torch.cuda.stream(torch.cuda.ExternalStream(fusion.stream(), device=device)):
        dist.all_reduce(t)
