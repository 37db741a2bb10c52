"""Multi-GPU host logic: independent filters shard across ranks with no hot-path communication; the only collective is
the final all-reduce of the statistics vector of fbus_stats."""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous index range [lo, hi) of the filters owned by `rank` (sizes differ by at most one)"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def combine_stats(vec, dist=None):
    """vec: torch tensor of FBUS_NSTATS doubles (fbus_stats layout): entries 0..4 are sums, entry 5 is a maximum.
    In-place all-reduce over the default process group (NCCL on GPUs, gloo in the CPU tests); returns vec."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return vec
    sums = vec[:5].clone()
    mx = vec[5:6].clone()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    vec[:5] = sums
    vec[5:6] = mx
    return vec


def summarize_stats(vec) -> dict:
    import math
    n_ok = max(float(vec[3]), 1.0)
    return {"rmse_pos_m": math.sqrt(float(vec[0]) / n_ok), "rmse_att_rad": math.sqrt(float(vec[1]) / n_ok),
            "nees_pose_mean_6dof": float(vec[2]) / n_ok, "filters_finite": int(vec[3]), "filters_nonfinite": int(vec[4]),
            "max_pos_err_m": float(vec[5])}
