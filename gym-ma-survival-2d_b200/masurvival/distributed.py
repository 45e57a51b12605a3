"""Multi-GPU host logic.  Environments are independent, so the batch shards
trivially: one process per GPU, each with its own handle over a contiguous
range of global env ids, and NO collective on the step path.
The only cross-rank operation is the optional sum of `flush_stats()`."""
import os


def rank_world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))


def shard(total_envs, rank, world):
    """Contiguous shard [first, first+count) of `total_envs` for `rank`.
    `first` is passed as `env_offset` so that every env keeps the Philox
    stream of its GLOBAL id whatever the number of GPUs."""
    base, rem = divmod(int(total_envs), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def all_reduce_stats(stats, group=None):
    """Sum a flush_stats() dict over ranks (works with gloo or nccl)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(stats)
    keys = sorted(stats)
    dev = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = {}
    for k, v in zip(keys, t.tolist()):
        out[k] = v if k.startswith('reward') else int(round(v))
    return out
