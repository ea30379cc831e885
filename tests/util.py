"""Shared helpers for the test-suite: seeded synthetic sequence generators and comparisons."""
from __future__ import annotations

import numpy as np

ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
NOISE = np.frombuffer(b"NnRYKMSWBDHV-*acgtUu\x00\x01\x02\x03 \n", dtype=np.uint8)


def random_batch(rng: np.random.Generator, lengths, noise: float = 0.0, n_runs: float = 0.0):
    """bases/offsets for sequences of the given lengths; `noise` = per-base probability of a byte from
    NOISE (ambiguity codes, lower case, U, raw 0..3 codes, junk); n_runs = per-sequence probability of
    an N run of random length."""
    lengths = np.asarray(lengths, dtype=np.uint64)
    offsets = np.zeros(len(lengths) + 1, dtype=np.uint64)
    np.cumsum(lengths, out=offsets[1:])
    total = int(offsets[-1])
    bases = ALPHA[rng.integers(0, 4, size=total)].copy()
    if noise > 0 and total:
        m = rng.random(total) < noise
        bases[m] = NOISE[rng.integers(0, len(NOISE), size=int(m.sum()))]
    if n_runs > 0:
        for i in np.nonzero(rng.random(len(lengths)) < n_runs)[0]:
            L = int(lengths[i])
            if L == 0:
                continue
            a = int(rng.integers(0, L))
            b = min(L, a + int(rng.integers(1, max(2, L // 3))))
            bases[int(offsets[i]) + a:int(offsets[i]) + b] = ord("N")
    return bases, offsets


def assert_rows_equal(got: np.ndarray, want_f64: np.ndarray, dtype, what=""):
    """Bit-exact comparison against the f64 oracle rows: integers for u32, f64 bits for f64 and
    (float)f64 for f32 (tolerance stated by north_star: 1 ulp f32; we hold 0 ulp)."""
    dtype = np.dtype(dtype)
    assert got.shape == want_f64.shape, (got.shape, want_f64.shape)
    if dtype == np.uint32:
        want = want_f64.astype(np.uint32)
    elif dtype == np.float32:
        want = want_f64.astype(np.float32)
    else:
        want = want_f64
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        i, j = bad[0]
        raise AssertionError(f"{what}: {len(bad)} mismatches, first at row {i} col {j}: "
                             f"got {got[i, j]!r} want {want[i, j]!r}")
