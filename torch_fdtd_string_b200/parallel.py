"""Multi-GPU plumbing for the stepper: one process per GPU, whole reference batches ("groups") per rank,
no collective on the data path (strings of different groups are independent).  torch.distributed
is used only for the barrier and for the max-over-ranks time."""
import torch


def rank_batches(n_batches, world_size, rank):
    """Batches (reference: one iteration of the loop at src/task/simulate.py:272) handled by `rank`:
    contiguous, sizes differ by at most one, every batch exactly once."""
    base, rem = divmod(n_batches, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (identity without an initialised process group)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def sum_over_ranks(value, device=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t)
