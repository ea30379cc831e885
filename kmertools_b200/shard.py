"""Multi-GPU sharding of a batch: contiguous read ranges balanced by bases, one range per rank.

Rows are independent (composition/src/oligo.rs:231-259 touches one sequence at a time), so every rank
computes its own rows and there is no collective on the data path.  The only collective
a launcher may want is an all_gather of the (lo, hi) ranges for bookkeeping.
"""
from __future__ import annotations

import numpy as np


def shard_by_bases(offsets: np.ndarray, world: int, rank: int) -> tuple[int, int]:
    """Sequence range [lo, hi) of `rank`: cut points are where the prefix sum of lengths crosses
    rank/world of the total, so ragged batches (contigs) are balanced by work, not by count."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    if world <= 1:
        return 0, n
    base0, total = int(offsets[0]), int(offsets[-1]) - int(offsets[0])
    if total == 0:  # all-empty batch: split by count
        return (n * rank) // world, (n * (rank + 1)) // world
    cuts = [0]
    for r in range(1, world):
        target = base0 + (total * r) // world
        cuts.append(int(np.searchsorted(offsets, np.uint64(target), side="left")))
    cuts.append(n)
    for i in range(1, len(cuts)):  # keep the cut points monotone
        cuts[i] = max(cuts[i], cuts[i - 1])
    return min(cuts[rank], n), min(cuts[rank + 1], n)


def local_batch(bases: np.ndarray, offsets: np.ndarray, world: int, rank: int):
    """(bases slice, rebased offsets, lo, hi) for `rank`."""
    lo, hi = shard_by_bases(offsets, world, rank)
    b0, b1 = int(offsets[lo]), int(offsets[hi])
    return bases[b0:b1], (offsets[lo:hi + 1] - offsets[lo]).astype(np.uint64), lo, hi
