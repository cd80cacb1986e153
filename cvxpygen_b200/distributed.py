"""Multi-GPU plumbing of the batched solve: instances never interact, so the batch is cut
into contiguous per-rank shards and NO data-path collective exists.  The only collective is an optional
one-off broadcast of the family constants (so every rank provably runs the same factor) and a tiny reduction of
counters for reporting.  One process per GPU, torch.distributed (nccl on GPUs, gloo in the CPU tests)."""
from typing import Dict, Tuple

import numpy as np


def shard_bounds(B: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of a batch of B instances; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError('rank out of range')
    base, rem = divmod(B, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_constants(blob: bytes, src: int = 0, device=None) -> bytes:
    """One broadcast of the constants blob from `src` (NCCL over NVLink on GPUs).  Every rank passes its own
    locally generated blob (used for the length) and receives rank src's bytes."""
    import torch
    import torch.distributed as dist
    t = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def reduce_counters(counters: Dict[str, float], device=None) -> Dict[str, float]:
    """Sum `n_*`/`sum_*` entries and max `max_*` entries over ranks (solved count, iteration sum, slowest time)."""
    import torch
    import torch.distributed as dist
    keys = sorted(counters)
    sums = torch.tensor([counters[k] for k in keys if not k.startswith('max_')], dtype=torch.float64, device=device)
    maxs = torch.tensor([counters[k] for k in keys if k.startswith('max_')] or [0.0], dtype=torch.float64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    out, i, j = {}, 0, 0
    for k in keys:
        if k.startswith('max_'):
            out[k] = float(maxs[j]); j += 1
        else:
            out[k] = float(sums[i]); i += 1
    return out


def shard_params(params: np.ndarray, world_size: int, rank: int) -> np.ndarray:
    lo, hi = shard_bounds(params.shape[0], world_size, rank)
    return params[lo:hi]
