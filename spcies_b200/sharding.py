"""Batch sharding across the GPUs of one box.

Instances are independent, so a batch is cut into contiguous slices, one per device / rank, and no collective
touches the data path.  The only communication is the reduction of *measurements* (max elapsed time, summed
counters) when one process per GPU is used (``bench.py`` under torchrun); the C ABI's own multi-device mode
(``spcies_batch_opts.n_devices``) uses one host thread per device and needs none.
"""
from __future__ import annotations


def shard_bounds(B: int, world: int):
    """Contiguous slices ``[lo, hi)`` of ``ceil(B / world)`` instances -- the same rule as Runtime::run()."""
    per = (B + world - 1) // world if world > 0 else B
    out = []
    for r in range(world):
        lo = min(B, r * per)
        out.append((lo, min(B, lo + per)))
    return out


def reduce_scalar(x: float, op: str = 'max', device=None) -> float:
    """All-reduce of one scalar over the default process group (nccl on GPUs, gloo on CPU); identity if not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op={'max': dist.ReduceOp.MAX, 'sum': dist.ReduceOp.SUM, 'min': dist.ReduceOp.MIN}[op])
    return float(t.item())
