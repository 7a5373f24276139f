"""Host-side sharding of independent alignment units over ranks (one process per GPU).

Batches of pairs and MSA jobs are independent units: rank r of W takes units r, r+W, r+2W, ...
(round-robin) and there is no collective on the data path; results are gathered
once at the end.  One long pair is different: it is cut into column slices (api.slice_columns)
and the ranks only exchange 64-byte IPC handles before the first run.

The functions take a `torch.distributed`-like object so that the same code runs under NCCL on
the GPU box and under gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_indices(nunits: int, rank: int, world: int) -> np.ndarray:
    """Indices of the units rank `rank` of `world` owns (round-robin)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return np.arange(rank, max(nunits, 0), world, dtype=np.int64)


def gather_results(dist, local_idx: np.ndarray, local_vals: np.ndarray, nunits: int, world: int) -> np.ndarray | None:
    """All ranks call this with their own indices/values (int32 [k, ncols]); rank 0 gets the
    [nunits, ncols] table in unit order, the others None.  One gather_object at job end."""
    payload = (np.asarray(local_idx, dtype=np.int64), np.asarray(local_vals, dtype=np.int32))
    if world == 1:
        parts = [payload]
    else:
        parts = [None] * world if dist.get_rank() == 0 else None
        dist.gather_object(payload, parts, dst=0)
        if dist.get_rank() != 0:
            return None
    ncols = payload[1].shape[1] if payload[1].ndim == 2 else 1
    out = np.zeros((nunits, ncols), dtype=np.int32)
    seen = np.zeros(nunits, dtype=bool)
    for idx, vals in parts:
        if len(idx):
            out[idx] = vals.reshape(len(idx), ncols)
            seen[idx] = True
    if not seen.all():
        raise RuntimeError("units missing from the gather: %d" % int((~seen).sum()))
    return out


def exchange_inbox_handles(dist, my_handle: bytes, world: int) -> list:
    """Every rank contributes the IPC handle of its left-boundary inbox (b"" when it has none);
    returns the list indexed by rank (rank r connects to entry r+1)."""
    if world == 1:
        return [my_handle]
    handles = [None] * world
    dist.all_gather_object(handles, my_handle)
    return handles
