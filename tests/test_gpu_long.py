"""Parity of the second-generation CTA-per-sequence kernel (csrc/long_kernel.cuh) with the CPU oracle:
MODE_K7 (k = 7: middle-base-first 16-bit keys, conflict-free scheduled write-out, bulk-copy rows) and MODE_FWD
(3 <= k <= 5: forward-code histogram folded at write-out).  Needs a B200: -m gpu."""
import numpy as np
import pytest

from tests.test_gpu_parity import check, comp
from tests.util import random_batch

pytestmark = pytest.mark.gpu

from kmertools_b200._lib import NORM_CLI, NORM_COUNTS, NORM_PY  # noqa: E402

EDGE_LENGTHS = [0, 1, 6, 7, 8, 15, 16, 17, 22, 23, 31, 32, 33, 150, 511, 512, 513, 527, 528, 1000, 4096, 10_000, 16_384,
                16_385, 70_001]


@pytest.mark.parametrize("dtype,norm", [(np.uint32, NORM_COUNTS), (np.float32, NORM_CLI), (np.float32, NORM_PY),
                                        (np.float32, NORM_COUNTS)])
def test_k7_edge_lengths(dtype, norm):
    rng = np.random.default_rng(7)
    lengths = np.array(EDGE_LENGTHS * 3)
    rng.shuffle(lengths)
    for noise, runs in ((0.0, 0.0), (0.01, 0.3), (0.3, 0.0)):
        bases, offsets = random_batch(rng, lengths, noise=noise, n_runs=runs)
        check(7, bases, offsets, norm_mode=norm, dtype=dtype, what=f"k7 noise={noise}")


def test_k7_matches_first_generation_kernel():
    rng = np.random.default_rng(8)
    bases, offsets = random_batch(rng, rng.integers(0, 30_000, size=300), noise=0.002, n_runs=0.2)
    a = check(7, bases, offsets, dtype=np.float32, what="k7 new")
    b = check(7, bases, offsets, dtype=np.float32, what="k7 old", k7_mid=0)
    assert np.array_equal(a, b)


def test_k7_many_short_reads_and_unaligned_starts():
    rng = np.random.default_rng(9)
    bases, offsets = random_batch(rng, rng.integers(100, 260, size=5000), noise=0.001)
    check(7, bases, offsets, dtype=np.float32, what="k7 short reads")
    check(7, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what="k7 short reads counts")


def test_k7_low_complexity_and_big_counts():
    """Homopolymers / dinucleotide repeats: one bin takes every window (same-address atomics); a sequence with more
    than 2^23 windows leaves the magic-constant float conversion."""
    seqs = [b"A" * 100_000, b"AC" * 40_000, b"ACG" * 30_000, b"T" * 9_000_000, b"ACGTTGCA" * 5_000]
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    check(7, bases, offsets, dtype=np.float32, what="k7 low complexity f32")
    check(7, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what="k7 low complexity u32")
    check(7, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.float32, what="k7 low complexity f32 counts")


@pytest.mark.parametrize("k", [3, 4, 5])
def test_forward_fold_all_lengths(k):
    """MODE_FWD forced on for every length (fwd_min_len = 0): short-kernel rejects, tiny and long sequences."""
    rng = np.random.default_rng(20 + k)
    lengths = np.array(EDGE_LENGTHS * 2 + [200_000])
    rng.shuffle(lengths)
    for noise, runs in ((0.0, 0.0), (0.02, 0.3)):
        bases, offsets = random_batch(rng, lengths, noise=noise, n_runs=runs)
        for dtype, norm in ((np.uint32, NORM_COUNTS), (np.float32, NORM_CLI), (np.float32, NORM_PY)):
            a = check(k, bases, offsets, norm_mode=norm, dtype=dtype, what=f"fwd k{k}", fwd_min_len=0)
            b = check(k, bases, offsets, norm_mode=norm, dtype=dtype, what=f"seq k{k}", fwd_fold=0)
            assert np.array_equal(a, b)


@pytest.mark.parametrize("k", [4, 5])
def test_forward_fold_contigs_default_dispatch(k):
    """Long ragged contigs with N runs, IUPAC codes and lower case take MODE_FWD by default (mean length >= 1024)."""
    rng = np.random.default_rng(40 + k)
    lengths = np.exp(rng.uniform(np.log(1e3), np.log(3e5), size=60)).astype(np.int64)
    bases, offsets = random_batch(rng, lengths, noise=0.003, n_runs=0.8)
    check(k, bases, offsets, dtype=np.float32, what=f"contigs k{k}")
    check(k, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"contigs k{k} counts")
    seqs = [b"A" * 300_000, b"GC" * 100_000, b"N" * 5000 + b"ACGT" * 1000 + b"N" * 17]
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    check(k, bases, offsets, dtype=np.float32, what=f"low complexity k{k}")


@pytest.mark.parametrize("k", [3, 4, 5])
def test_forward_fold_replicated_bins(k):
    """Long contigs (mean length >= 32 kbp) at k <= 5 count into lane-private replicas of the bins (32 at k <= 4, 8 at
    k = 5) that are summed at write-out; same rows as without replicas and as the oracle, palindromes included."""
    rng = np.random.default_rng(60 + k)
    lengths = np.r_[np.exp(rng.uniform(np.log(2e4), np.log(4e5), size=24)).astype(np.int64), [0, 3, k - 1, k, 700, 33_000]]
    rng.shuffle(lengths)
    bases, offsets = random_batch(rng, lengths, noise=0.004, n_runs=0.8)
    for dtype, norm in ((np.uint32, NORM_COUNTS), (np.float32, NORM_CLI), (np.float32, NORM_PY)):
        a = check(k, bases, offsets, norm_mode=norm, dtype=dtype, what=f"replicas k{k}", fwd_replicas=1)
        b = check(k, bases, offsets, norm_mode=norm, dtype=dtype, what=f"plain k{k}", fwd_replicas=0)
        assert np.array_equal(a, b)
    seqs = [b"A" * 500_000, b"ACGT" * 100_000, b"GC" * 200_000 + b"N" * 1000 + b"TTAA" * 50_000]   # hot bins, palindromes
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    check(k, bases, offsets, dtype=np.float32, what=f"replicas low complexity k{k}")
    check(k, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"replicas low complexity k{k} counts")


@pytest.mark.parametrize("k", [4, 5])
def test_longest_first_order_for_many_contigs(k):
    """More long contigs than CTAs: the work queue is re-ordered by length class (order_count_kernel /
    order_scatter_kernel).  Rows must not depend on the order; short members of the batch ride along."""
    rng = np.random.default_rng(80 + k)
    lengths = np.exp(rng.uniform(np.log(2e3), np.log(2.5e5), size=1500)).astype(np.int64)
    lengths[::97] = rng.integers(0, 200, size=len(lengths[::97]))   # reads short enough for short_kernel's groups
    lengths[5] = 0
    bases, offsets = random_batch(rng, lengths, noise=0.002, n_runs=0.5)
    assert bases.size // len(lengths) >= 32768
    a = check(k, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"longest first k{k}", longest_first=1)
    b = check(k, bases, offsets, norm_mode=NORM_COUNTS, dtype=np.uint32, what=f"input order k{k}", longest_first=0)
    assert np.array_equal(a, b)
    check(k, bases, offsets, norm_mode=NORM_CLI, dtype=np.float32, what=f"longest first k{k} f32")
