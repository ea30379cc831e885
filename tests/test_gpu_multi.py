"""The sharded multi-GPU entry point (csrc/multi.cu) on real devices: one call, contiguous ranges balanced by
bases, every device writes its own rows, result identical to the oracle and to a single device.  Runs on however many
GPUs the box has (the 2-device cases are skipped on a 1-GPU box).  Needs a B200: -m gpu."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_rows_equal, random_batch

pytestmark = pytest.mark.gpu

from kmertools_b200 import MultiOligoComputer, OligoComputer, _lib, shard_bounds  # noqa: E402
from kmertools_b200._lib import NORM_CLI, NORM_COUNTS, NORM_PY  # noqa: E402


def ndev():
    return _lib.load().ktb_device_count()


def _batch(seed=1):
    rng = np.random.default_rng(seed)
    lengths = np.r_[rng.integers(0, 400, size=3000), [50_000, 0, 0, 7, 120_000], rng.integers(100, 260, size=2000)]
    return random_batch(rng, lengths, noise=0.01, n_runs=0.05)


@pytest.mark.parametrize("k,dtype,norm", [(5, np.float32, NORM_CLI), (4, np.uint32, NORM_COUNTS), (7, np.float32, NORM_PY),
                                          (9, np.uint32, NORM_COUNTS), (3, np.float64, NORM_CLI)])
def test_multi_matches_oracle_on_all_devices(k, dtype, norm):
    bases, offsets = _batch(k)
    if k == 9:
        bases, offsets = random_batch(np.random.default_rng(9), [3000, 0, 40_000, 12, 9000, 70_000], noise=0.002)
    mc = MultiOligoComputer(k)            # every visible GPU
    assert mc.ndev == ndev()
    n = len(offsets) - 1
    totals = np.zeros(n, dtype=np.uint64)
    got = mc.vectorise_packed(bases, offsets, norm_mode=norm, dtype=dtype, totals=totals)
    want, wt = O.vectorise_batch(bases, offsets, k, True, norm)
    assert np.array_equal(totals, wt)
    assert_rows_equal(got, want, dtype, f"multi k{k}")
    # the devices' row ranges are the library's partition, contiguous and complete
    b = shard_bounds(offsets, mc.ndev)
    rows = [(mc.device_stats(i)["first_row"], mc.device_stats(i)["end_row"]) for i in range(mc.ndev)]
    assert rows == [(int(b[i]), int(b[i + 1])) for i in range(mc.ndev)]
    assert sum(mc.device_stats(i)["launches"] for i in range(mc.ndev)) > 0
    mc.close()


def test_multi_numa_local_rows_and_reference_semantics():
    bases, offsets = _batch(3)
    mc = MultiOligoComputer(5)
    rows = mc.alloc_rows(offsets, dtype=np.float32)      # slab of device i on that device's NUMA node, page-locked
    got = mc.vectorise_packed(bases, offsets, out=rows.array)
    want, _ = O.vectorise_batch(bases, offsets, 5, True, NORM_CLI)
    assert_rows_equal(got, want, np.float32, "numa rows")
    seqs = [bytes(bases[int(offsets[i]):int(offsets[i + 1])]).decode("latin-1") for i in range(40)]
    a = mc.vectorise_batch(seqs)
    oc = OligoComputer(5)
    assert a == oc.vectorise_batch(seqs)                 # reference semantics, same as one device
    oc.close()
    rows.free()
    mc.close()
    assert isinstance(_lib.load().ktb_device_numa_node(0), int)


@pytest.mark.skipif(ndev() < 2, reason="needs two GPUs")
def test_two_devices_split_and_single_device_agree():
    bases, offsets = _batch(4)
    two = MultiOligoComputer(5, devices=[1, 0])          # explicit list, order = order of the ranges
    one = MultiOligoComputer(5, devices=[0])
    a = two.vectorise_packed(bases, offsets, dtype=np.float32)
    b = one.vectorise_packed(bases, offsets, dtype=np.float32)
    assert np.array_equal(a, b)
    s0, s1 = two.device_stats(0), two.device_stats(1)
    assert s0["end_row"] == s1["first_row"] and s0["first_row"] == 0 and s1["end_row"] == len(offsets) - 1
    assert s0["launches"] > 0 and s1["launches"] > 0 and s0["d2h_bytes"] > 0 and s1["d2h_bytes"] > 0
    # balanced by bases within one sequence
    nb0 = int(offsets[s0["end_row"]]) - int(offsets[0])
    assert abs(nb0 - int(offsets[-1]) / 2) <= 120_000
    two.close(); one.close()


def test_current_device_is_restored():
    """entry points must not leave the caller's current device switched."""
    import torch
    if ndev() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    oc = OligoComputer(4, device=1)
    oc.vectorise_batch(["ACGTACGTTTGA"])
    assert torch.cuda.current_device() == 0
    x = torch.ones(4, device="cuda")
    assert x.device.index == 0
    oc.close()


def _rank_worker(rank: int, world: int, port: int, ret):
    """One process per rank as under torchrun: every rank computes ITS contiguous range on a GPU (rank % #GPUs, so two
    ranks share the device of a 1-GPU box), ranks exchange only bookkeeping (gloo)."""
    import os
    import torch.distributed as dist
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kmertools_b200.shard import local_batch
        rng = np.random.default_rng(123)   # the same batch on every rank (as a shared input file would be)
        lengths = np.r_[rng.integers(0, 400, size=300), [50_000, 0, 0, 7], rng.integers(100, 3000, size=100)]
        bases, offsets = random_batch(rng, lengths, noise=0.01)
        lb, lo_offs, lo, hi = local_batch(bases, offsets, world, rank)
        b = shard_bounds(offsets, world)
        assert (lo, hi) == (int(b[rank]), int(b[rank + 1]))          # python sharder == library partition
        oc = OligoComputer(5, device=rank % ndev())
        rows = oc.vectorise_packed(lb, lo_offs, norm_mode=NORM_CLI, dtype=np.float32)
        assert oc.stats()["launches"] > 0
        full, _ = O.vectorise_batch(bases, offsets, 5, True, NORM_CLI)
        assert_rows_equal(rows, full[lo:hi], np.float32, f"rank {rank}")
        t = torch.tensor([lo, hi], dtype=torch.int64)
        got = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(got, t)
        ranges = [tuple(int(x) for x in g) for g in got]
        assert ranges[0][0] == 0 and ranges[-1][1] == len(lengths) and all(a[1] == c[0] for a, c in zip(ranges, ranges[1:]))
        oc.close()
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_ranks_compute_their_shards_on_gpus():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_rank_worker, args=(r, world, port, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=300)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(ret) == {0: True, 1: True}
