"""world_size-2 checks of the N>1 path on CPU (gloo): ranks agree on a partition of the batch that covers
every sequence exactly once, balanced by bases; the only collective is bookkeeping."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kmertools_b200.shard import local_batch, shard_by_bases
from oracle import oracle as O
from tests.util import random_batch


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)  # same batch on every rank (as a shared input file would be)
        lengths = np.r_[rng.integers(0, 400, size=300), [50_000, 0, 0, 7]]
        bases, offsets = random_batch(rng, lengths, noise=0.01)
        lb, lo_offs, lo, hi = local_batch(bases, offsets, world, rank)
        # each rank computes ITS rows (the oracle stands in for the GPU here; this test is about sharding)
        rows, _ = O.vectorise_batch(lb, lo_offs, 4, True, 1)
        # bookkeeping collective: gather the ranges
        rng_t = torch.tensor([lo, hi, int(lo_offs[-1])], dtype=torch.int64)
        gathered = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, rng_t)
        ranges = [tuple(int(x) for x in g) for g in gathered]
        # partition: contiguous, complete, non-overlapping
        assert ranges[0][0] == 0 and ranges[-1][1] == len(lengths)
        for a, b in zip(ranges, ranges[1:]):
            assert a[1] == b[0]
        assert sum(r[2] for r in ranges) == int(offsets[-1])
        # my rows equal the corresponding rows of the unsharded computation
        full, _ = O.vectorise_batch(bases, offsets, 4, True, 1)
        assert np.array_equal(rows, full[lo:hi])
        # balanced by bases within one sequence length of the ideal split
        ideal = int(offsets[-1]) / world
        assert abs(ranges[rank][2] - ideal) <= lengths.max()
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(ret) == {0: True, 1: True}


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_ranges_cover_everything(world):
    rng = np.random.default_rng(world)
    for lengths in (rng.integers(0, 1000, size=97), np.zeros(10, dtype=np.int64), np.array([5]),
                    np.r_[np.full(50, 150), [10**6], np.full(50, 150)]):
        offsets = np.zeros(len(lengths) + 1, dtype=np.uint64)
        np.cumsum(lengths, out=offsets[1:])
        prev = 0
        for r in range(world):
            lo, hi = shard_by_bases(offsets, world, r)
            assert lo == prev and hi >= lo
            prev = hi
        assert prev == len(lengths)
