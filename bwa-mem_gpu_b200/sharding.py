"""Read-batch sharding across ranks: units are reads, contiguous blocks per
rank, FMD index replicated on every GPU, no data-path collective.  The only communication is the
optional final gather of the fixed-size per-read records onto rank 0 (torch.distributed; NCCL on
GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous block [lo, hi) of rank `rank`; blocks differ in size by at most one unit"""
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_records(local: np.ndarray, n_total: int, dist, rank: int, world: int, device: str = "cpu"):
    """gather per-read records (structured numpy array) of every rank's block onto rank 0, in read order"""
    import torch
    if world == 1:
        return local
    raw = torch.from_numpy(local.view(np.uint8).reshape(-1).copy()).to(device)
    sizes = [(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0]) * local.dtype.itemsize for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=torch.uint8, device=device)
    buf[:raw.numel()] = raw
    outs = [torch.zeros(pad, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, outs, dst=0)
    if rank != 0:
        return None
    parts = [outs[r][:sizes[r]].cpu().numpy().view(local.dtype) for r in range(world)]
    return np.concatenate(parts)
