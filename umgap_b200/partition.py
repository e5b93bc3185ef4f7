"""Read partitioning for the multi-GPU (replicated-index) mode.

Reads are independent units, so N GPUs run N replicas of the table and taxonomy and each rank
classifies a contiguous slice of the read groups; there is no collective on the data path.  The
only exchange is the gather of the per-group results (and the max-over-ranks of the timings),
which goes through ``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np


def split_groups(group_off: np.ndarray, read_off: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous [g0, g1) group ranges, one per rank, balanced by nucleotide count; a group (the
    reads `uniq` joins) is never split."""
    ngroups = len(group_off) - 1
    if world <= 0:
        raise ValueError("world must be positive")
    nt_before = np.asarray(read_off, dtype=np.uint64)[np.asarray(group_off, dtype=np.uint64)]
    total = int(nt_before[-1]) if ngroups else 0
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        g = int(np.searchsorted(nt_before, target, side="left"))
        bounds.append(min(max(g, bounds[-1]), ngroups))
    bounds.append(ngroups)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def slice_batch(nt: np.ndarray, read_off: np.ndarray, group_off: np.ndarray, g0: int, g1: int):
    """The sub-batch holding groups [g0, g1) with offsets rebased to zero."""
    r0, r1 = int(group_off[g0]), int(group_off[g1])
    n0, n1 = int(read_off[r0]), int(read_off[r1])
    return (nt[n0:n1], (read_off[r0:r1 + 1] - read_off[r0]).astype(np.uint64),
            (group_off[g0:g1 + 1] - group_off[g0]).astype(np.uint64))


def classify_partitioned(classify: Callable[[np.ndarray, np.ndarray, np.ndarray], np.ndarray], nt: np.ndarray,
                         read_off: np.ndarray, group_off: np.ndarray, rank: int, world: int, dist=None) -> np.ndarray:
    """Every rank holds the whole batch, classifies its slice with `classify(nt, read_off, group_off)`
    (on its own GPU replica) and all ranks end up with the per-group results in input order."""
    parts = split_groups(group_off, read_off, world)
    g0, g1 = parts[rank]
    local = np.asarray(classify(*slice_batch(nt, read_off, group_off, g0, g1)), dtype=np.uint32)
    if local.shape[0] != g1 - g0:
        raise ValueError("classify returned a wrong number of results")
    if world == 1 or dist is None:
        return local
    import torch
    longest = max(b - a for a, b in parts)
    buf = torch.zeros(longest, dtype=torch.int64)
    buf[: g1 - g0] = torch.from_numpy(local.astype(np.int64))
    out = [torch.zeros(longest, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([out[r][: b - a].numpy() for r, (a, b) in enumerate(parts)]).astype(np.uint32)
