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
