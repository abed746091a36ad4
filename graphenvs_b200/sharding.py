"""Multi-GPU plumbing: the env batch shards into independent slices, one per rank.
No data-path collective exists; the only exchange is a SUM all-reduce of the episode statistics."""
import torch


def rank_slice(total_envs, rank, world):
    """(first global env id, count) of `rank`'s slice: env i -> rank i // ceil(total/world)."""
    per = (total_envs + world - 1) // world
    lo = min(rank * per, total_envs)
    return lo, max(0, min(per, total_envs - lo))


def reduce_stats(stats, group=None):
    """In-place SUM all-reduce of a statistics vector (episodes, solved, sum reward, sum cost, ...).
    NCCL when `stats` lives on a GPU, gloo on CPU; a no-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def max_over_ranks(value, device=None, group=None):
    """Max of a python float over ranks (device timings are reported as the max over ranks)."""
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
