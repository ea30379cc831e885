"""Parity of the bucket-then-count path (csrc/bucket_kernels.cuh: rows larger than shared memory, canonical k = 9, 10 and
raw k = 8..10) with the CPU oracle, including ALL 2,000 sequences of BASELINE config 5 (ii).  Needs a B200: -m gpu."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.test_gpu_parity import check
from tests.util import assert_rows_equal, random_batch

pytestmark = pytest.mark.gpu

from kmertools_b200._lib import NORM_CLI, NORM_COUNTS, NORM_PY  # noqa: E402

TILE = 248 * 16   # bases per tile of bucket_kernel


@pytest.mark.parametrize("k,mins", [(9, True), (10, True), (8, False), (9, False)])
def test_bucket_path_ragged(k, mins):
    rng = np.random.default_rng(900 + k + 10 * mins)
    lengths = np.r_[rng.integers(0, 4000, size=20), [0, 1, k - 1, k, k + 1, 15, 16, 17, 495, 496, 497, 511, 512, 513],
                    [TILE - 1, TILE, TILE + 1, TILE + k - 1, 2 * TILE - 3, 3 * TILE + 500], [120_000]]
    rng.shuffle(lengths)
    for noise, runs in ((0.0, 0.0), (0.003, 0.3)):
        bases, offsets = random_batch(rng, lengths, noise=noise, n_runs=runs)
        a = check(k, bases, offsets, mins=mins, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"bucket k{k} u32")
        b = check(k, bases, offsets, mins=mins, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"wave k{k} u32", bucket=0)
        assert np.array_equal(a, b)
        check(k, bases, offsets, mins=mins, norm_mode=NORM_CLI, dtype=np.float32, what=f"bucket k{k} f32")
        check(k, bases, offsets, mins=mins, norm_mode=NORM_PY, dtype=np.float32, what=f"bucket k{k} f32 py")
    if k == 9:
        check(k, bases, offsets, mins=mins, norm_mode=NORM_CLI, dtype=np.float64, what=f"bucket k{k} f64")
        check(k, bases, offsets, mins=mins, norm_mode=NORM_COUNTS, dtype=np.float64, what=f"bucket k{k} f64 counts")


@pytest.mark.parametrize("log2_seg", [13, 14])
def test_bucket_segment_sizes(log2_seg):
    rng = np.random.default_rng(77 + log2_seg)
    lengths = np.r_[rng.integers(0, 50_000, size=12), [TILE * 2 + 7]]
    bases, offsets = random_batch(rng, lengths, noise=0.002, n_runs=0.2)
    check(9, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"k9 seg 2^{log2_seg}", bucket_log2_seg=log2_seg)
    check(9, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, what=f"k9 seg 2^{log2_seg} f32", bucket_log2_seg=log2_seg)
    check(8, bases, offsets, mins=False, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"raw k8 seg 2^{log2_seg}",
          bucket_log2_seg=log2_seg)
    check(10, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"k10 seg 2^{log2_seg} (2^13: falls back)",
          bucket_log2_seg=log2_seg)


def test_bucket_low_complexity_and_unaligned_starts():
    """Every window of a homopolymer lands in ONE segment run (a tile's run reaches its maximum length) and in one
    bin; sequences start at every 16-byte phase."""
    seqs = [b"A" * 200_000, b"G" * 7, b"AC" * 60_000, b"ACGTTGCAAC" * 9_000, b"T" * (TILE + 9), b"N" * 40_000 + b"ACGT" * 10,
            b"ACGTN" * 30_000]
    seqs += [b"GATTACA" * (3 + i) for i in range(17)]
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    for k in (9, 10):
        check(k, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"k{k} low complexity")
        check(k, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, what=f"k{k} low complexity f32")


def test_bucket_one_long_contig():
    """A 3 Mbp contig = 757 tiles of one sequence: count_kernel walks its runs in three passes of 256 descriptors."""
    rng = np.random.default_rng(5)
    bases, offsets = random_batch(rng, [700, 3_000_000, 0, 2500], noise=0.0005, n_runs=0.5)
    check(9, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what="k9 long contig")
    check(10, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, what="k10 long contig f32")
    check(8, bases, offsets, mins=False, norm_mode=NORM_COUNTS, dtype=np.uint32, what="raw k8 long contig")


def test_config5ii_all_sequences_bit_exact():
    """config 5 (ii) — 2,000 x 100 kbp, k = 10 canonical u32 counts — every sequence against the oracle
    (the oracle runs in slices of 50 rows to bound host memory)."""
    import torch
    import bench
    from kmertools_b200 import OligoComputer
    spec = bench.WORKLOADS["reads100k_k10"]
    dev = torch.device("cuda", 0)
    bases, offsets = bench.make_workload_torch(spec, 1.0, dev)
    n, k = offsets.numel() - 1, spec["k"]
    assert n == 2000
    oc = OligoComputer(k)
    totals = torch.zeros(n, dtype=torch.int64, device=dev)
    counts = oc.vectorise_tensors(bases, offsets, norm_mode=0, dtype=torch.int32, totals=totals)
    torch.cuda.synchronize()
    hb = bases.cpu().numpy()
    ho = offsets.cpu().numpy().astype(np.uint64)
    ht = totals.cpu().numpy().astype(np.uint64)
    for a in range(0, n, 50):
        b = min(n, a + 50)
        so = ho[a:b + 1] - ho[a]
        want, wt = O.vectorise_batch(hb[int(ho[a]):int(ho[b])], so, k, True, 0)
        assert np.array_equal(ht[a:b], wt)
        assert_rows_equal(counts[a:b].cpu().numpy().view(np.uint32), want, np.uint32, f"config 5ii rows {a}..{b}")
    oc.close()


@pytest.mark.parametrize("waves", [1, 3, 8, 64])
def test_bucket_waves_agree(waves):
    """bucket_kernel(w+1) overlaps count_kernel(w) on two streams: any number of waves gives the same rows (wave
    boundaries fall on chunks of 8 sequences; the last wave is ragged)."""
    rng = np.random.default_rng(31)
    lengths = np.r_[rng.integers(0, 9000, size=300), [0, 8, 9, 10, 50_000], rng.integers(3000, 5000, size=133)]
    bases, offsets = random_batch(rng, lengths, noise=0.002, n_runs=0.1)
    check(9, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"k9 {waves} waves", bucket_waves=waves)
    check(9, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, what=f"k9 {waves} waves f32", bucket_waves=waves)
    check(8, bases, offsets, mins=False, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"raw k8 {waves} waves", bucket_waves=waves)
