"""Multi-GPU plumbing: instances are independent, so the batch shards by index with no data-path collective.
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used for exactly three things: the barrier around the
timed region, max-over-ranks of the device time, and the final sum of per-rank accepted/rejected counts."""
import torch


def shard_range(n_total, rank, world):
    """contiguous index range [lo, hi) of `n_total` instances owned by `rank`"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seed(seed, rank, per_rank):
    """instance i of the global batch is generated from seed + i on whichever rank owns it"""
    return seed + rank * per_rank


def reduce_timing_and_counts(dist, device, times_ms, counts):
    """times -> max over ranks, counts -> sum over ranks.  `dist` is torch.distributed or None (single process)."""
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    c = torch.tensor(list(counts), dtype=torch.int64, device=device)
    if dist is not None and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()], [int(v) for v in c.tolist()]
