"""Property-based parity (hypothesis): random k, alphabets (ACGT either case, U, raw 0..3 codes, IUPAC, junk, bytes >= 128),
ragged lengths around 0 / k-1 / k / chunk and step boundaries, canonical and raw mode, three output types, all three
normalisation modes — every example bit-exact against the CPU oracle.  Needs a B200: -m gpu."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as O
from tests.test_gpu_parity import comp
from tests.util import assert_rows_equal

pytestmark = pytest.mark.gpu

from kmertools_b200._lib import NORM_CLI, NORM_COUNTS, NORM_PY  # noqa: E402

ALPHABETS = {
    "acgt": b"ACGT",
    "mixed_case_u": b"ACGTacgtUu",
    "with_n": b"ACGTACGTACGTACGTN",
    "iupac": b"ACGTRYKMSWBDHVN",
    "raw_codes": bytes([0, 1, 2, 3]) + b"ACGT",
    "junk": b"ACGTacgt-* \n\t.@>+",
    "high_bytes": b"ACGT" + bytes([128, 200, 255, 0xC1, 0xE7]),
    "any_byte": bytes(range(256)),
}


@st.composite
def batches(draw):
    k = draw(st.integers(1, 10))
    alpha = ALPHABETS[draw(st.sampled_from(sorted(ALPHABETS)))]
    nseq = draw(st.integers(1, 12))
    edge = [0, 1, k - 1, k, k + 1, 15, 16, 17, 31, 32, 33, 255 + k, 254 + k, 511, 512, 513, 527]
    lengths = draw(st.lists(st.one_of(st.sampled_from(edge), st.integers(0, 700), st.integers(0, 5000)),
                            min_size=nseq, max_size=nseq))
    if k >= 9:   # rows of 0.5 - 2 MB: keep the oracle's output small
        lengths = lengths[:4]
    seed = draw(st.integers(0, 2 ** 32 - 1))
    rng = np.random.default_rng(seed)
    arr = np.frombuffer(alpha, dtype=np.uint8)
    seqs = [arr[rng.integers(0, len(arr), size=max(0, L))] for L in lengths]
    if draw(st.booleans()) and seqs:   # a homopolymer / low-complexity member
        i = draw(st.integers(0, len(seqs) - 1))
        seqs[i] = np.full(len(seqs[i]), ord(draw(st.sampled_from("ACGTN"))), dtype=np.uint8)
    mins = draw(st.booleans())
    if not mins and k > 8:
        k = 8          # raw rows of 4^k columns: keep them below 256 KB
        lengths = lengths[:4]
        seqs = seqs[:4]
    dtype, norm = draw(st.sampled_from([(np.uint32, NORM_COUNTS), (np.float32, NORM_CLI), (np.float32, NORM_PY),
                                        (np.float32, NORM_COUNTS), (np.float64, NORM_CLI), (np.float64, NORM_PY)]))
    return k, seqs, mins, dtype, norm


@settings(max_examples=250, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(batches())
def test_random_batches_match_the_oracle(example):
    k, seqs, mins, dtype, norm = example
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    bases = np.concatenate(seqs).astype(np.uint8) if seqs and offsets[-1] else np.zeros(0, dtype=np.uint8)
    oc = comp(k)
    totals = np.zeros(len(seqs), dtype=np.uint64)
    got = oc.vectorise_packed(bases, offsets, norm_mode=norm, mins=mins, dtype=dtype, totals=totals)
    want, wtot = O.vectorise_batch(bases, offsets, k, mins, norm)
    assert np.array_equal(totals, wtot)
    assert_rows_equal(got, want, dtype, f"k={k} mins={mins} norm={norm} dtype={np.dtype(dtype).name}")
