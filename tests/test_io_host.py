"""CPU-side tests of the host feeder (FASTA/FASTQ reader) and of the text formatter's rounding."""
import gzip

import numpy as np
import pytest

from kmertools_b200 import io as kio
from oracle import oracle as O


def _as_list(bases, offsets):
    return [bytes(bases[int(offsets[i]):int(offsets[i + 1])]) for i in range(len(offsets) - 1)]


@pytest.mark.parametrize("fname", ["reads.fa", "reads.fq", "reads.fq.gz"])
@pytest.mark.parametrize("sniff", [False, True])
def test_reader_matches_reference_fixtures(golden, fname, sniff):  # ktio/src/seq.rs:164-233
    bases, offsets = kio.read_fastx(golden / fname, sniff=sniff)
    want = [s for _, s in O.read_fastx(golden / fname)]
    assert _as_list(bases, offsets) == want
    assert len(want) == 2 and int(offsets[-1]) == 144


def test_reader_fasta_without_trailing_newline(tmp_path):  # load_fa_stdin_test
    p = tmp_path / "x.fa"
    p.write_bytes(b">Record_1\nACGTACGTACGT")
    assert _as_list(*kio.read_fastx(p)) == [b"ACGTACGTACGT"]


def test_reader_synthetic_variants(tmp_path):
    rng = np.random.default_rng(3)
    seqs = [bytes(rng.choice(list(b"ACGTN"), size=int(n)).astype(np.uint8)) for n in
            [0, 1, 59, 60, 61, 300, 5000, 70000, 17]]
    fa = b"".join(b">s%d some description\n" % i + b"\n".join(s[j:j + 60] for j in range(0, len(s), 60)) + b"\n"
                  for i, s in enumerate(seqs))
    fq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate(seqs) if s)
    (tmp_path / "a.fasta").write_bytes(fa)
    (tmp_path / "a.fna").write_bytes(fa.replace(b"\n", b"\r\n"))
    (tmp_path / "b.fastq").write_bytes(fq)
    with gzip.open(tmp_path / "b.fq.gz", "wb") as fh:
        fh.write(fq)
    with gzip.open(tmp_path / "a.fa.gz", "wb") as fh:
        fh.write(fa)
    assert _as_list(*kio.read_fastx(tmp_path / "a.fasta")) == seqs
    assert _as_list(*kio.read_fastx(tmp_path / "a.fna")) == seqs       # CRLF: trailing whitespace trimmed
    assert _as_list(*kio.read_fastx(tmp_path / "a.fa.gz")) == seqs
    assert _as_list(*kio.read_fastx(tmp_path / "b.fastq")) == [s for s in seqs if s]
    assert _as_list(*kio.read_fastx(tmp_path / "b.fq.gz")) == [s for s in seqs if s]
    assert _as_list(*kio.read_fastx(tmp_path / "b.fastq", sniff=True)) == [s for s in seqs if s]


def test_reader_errors(tmp_path):
    from kmertools_b200 import KtbError
    with pytest.raises(KtbError, match="Unable to open"):
        kio.read_fastx(tmp_path / "missing.fa")
    p = tmp_path / "x.txt"
    p.write_bytes(b">a\nACGT\n")
    with pytest.raises(KtbError, match="extension"):
        kio.read_fastx(p)
    assert _as_list(*kio.read_fastx(p, sniff=True)) == [b"ACGT"]
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b"ACGT\n>a\nAC\n")
    with pytest.raises(KtbError, match="Expected >"):
        kio.read_fastx(bad)


def test_format6_is_correctly_rounded_half_even():
    """format6 == C's "%.6f" (exact binary value, ties to even) == Rust's "{:.6}"."""
    cases = [0.0, 1.0, 1 / 128, 3 / 128, 5 / 128, 1 / 64, 1 / 2, 0.5 ** 20, 0.5 ** 21, 0.5 ** 30, 1e-7, 4.9999995e-7,
             5e-7, 5.0000005e-7, 0.9999995, 0.99999949999, 0.9999996, 1 / 3, 2 / 3, 1 / 146, 1 / 69, 1 / 26]
    rng = np.random.default_rng(0)
    T = rng.integers(1, 1 << 24, size=60000)
    c = (rng.random(60000) * (T + 1)).astype(np.int64)
    cases += list(np.minimum(c, T) / T)
    cases += [cc / t for t in range(1, 300) for cc in range(0, t + 1)]
    # exact ties at the 6th decimal are dyadic: k/2^j with a trailing ...5 at the 7th decimal
    cases += [k / 128 for k in range(129)] + [k / 2048 for k in range(0, 2049, 7)]
    for q in cases:
        assert kio.format6(float(q)) == "%.6f" % float(q), q
