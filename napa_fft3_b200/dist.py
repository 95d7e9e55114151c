"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only for control.

Batched transforms (BASELINE configs C2-C4) are independent, so the batch is cut into contiguous
per-rank slices and every rank runs the single-GPU path on its slice: no data-path collective
(SURVEY.md 8e).  `gather_timing` is the max-over-ranks reduction bench.py reports."""
from __future__ import annotations

import numpy as np


def shard_bounds(batch: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of `batch` items owned by `rank`; sizes differ by at most one and
    every item is owned exactly once (ragged batches included)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(x: np.ndarray, rank: int, world: int) -> np.ndarray:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def sharded_apply(fn, x: np.ndarray, rank: int, world: int, gather=None):
    """Run `fn` (e.g. napa.fft_batch) on this rank's rows.  `gather(local) -> list of per-rank
    arrays` is optional (tests use torch.distributed.all_gather_object; production keeps results
    sharded)."""
    local = fn(np.ascontiguousarray(shard_rows(x, rank, world)))
    if gather is None:
        return local
    return np.concatenate(gather(local), axis=0)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """bench.py timing rule: a multi-GPU number is the max over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
