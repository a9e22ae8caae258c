"""Image-level sharding for multi-GPU runs (one process per GPU, no collective on the hot path).

Reference equivalent: ``DistributedSampler(dataset_val, shuffle=False)`` (RV/main.py:239-241) -- but unlike the
sampler, nothing is padded or dropped: every image is processed exactly once and results are gathered on the host
after the timed region, then ordered by filename like ``SubmissionWriter.export`` (RV/utils/submission.py:41-43).
"""
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous [start, stop) slice of ``range(n_items)`` for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def batches(start, stop, batch_size):
    """Batch boundaries inside a shard; the ragged tail is a short last batch, never dropped."""
    return [(i, min(i + batch_size, stop)) for i in range(start, stop, batch_size)]


def gather_results(local, dst=0):
    """Gather per-rank result dicts ({filename: (quat, tvec, status)}) on ``dst`` after the timed region."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(sorted(local.items()))
    ws = dist.get_world_size()
    out = [None] * ws if dist.get_rank() == dst else None
    dist.gather_object(local, out, dst=dst)
    if dist.get_rank() != dst:
        return None
    merged = {}
    for part in out:
        merged.update(part)
    return dict(sorted(merged.items()))
