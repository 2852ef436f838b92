"""Multi-GPU sharding for the featExtract path.

Batch mode (BASELINE.json config 4): volumes are independent, so volume i goes to rank i % world and
there is NO data-path collective; the only communication is the final gather of the (small) feature
lists to rank 0, which restores input order.  One process per GPU, ``torch.distributed`` (NCCL on
GPUs, gloo in the CPU tests) for the plumbing.
"""
import numpy as np


def shard_indices(n_items, rank, world):
    """Indices of the items rank ``rank`` owns: round-robin, like the reference running one process
    per volume would be scheduled (no volume is split)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_items, world))


def extract_sharded(extract_fn, volumes, rank=0, world=1, group=None, dst=0):
    """Run ``extract_fn(volume) -> ndarray`` on this rank's shard of ``volumes`` and gather the
    per-volume results on ``dst`` in input order (other ranks get None).

    ``volumes`` may be a sequence or a callable ``i -> volume`` with ``len`` given by ``n_items``
    attribute; only the owned indices are touched, so every rank can hold just its own data.
    """
    n = len(volumes)
    mine = shard_indices(n, rank, world)
    local = [(i, extract_fn(volumes[i])) for i in mine]
    if world == 1:
        out = [None] * n
        for i, f in local:
            out[i] = f
        return out
    import torch.distributed as dist
    gathered = [None] * world if rank == dst else None
    dist.gather_object(local, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = [None] * n
    for part in gathered:
        for i, f in part:
            out[i] = f
    assert all(o is not None for o in out)
    return out


def max_over_ranks(value, device=None, group=None):
    """Max of a python float over all ranks (timing is reported as the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
