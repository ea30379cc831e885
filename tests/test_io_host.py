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


# ---------------------------------------------------------------------------------------------------------------
# Parser rules beyond the reference's own fixtures.  The reference reads records with rust-bio 2.3.0
# (bio::io::fasta::Reader::read / bio::io::fastq::Reader::read, called through Records::next and unwrap()ed in
# ktio/src/seq.rs:97-139) and inflates ".gz" with flate2::read::GzDecoder (seq.rs:149).  Neither crate is in
# /root/reference and Rust cannot run here, so each case states the rule of the published source it restates.
def _write(tmp_path, name, data):
    p = tmp_path / name
    p.write_bytes(data)
    return p


def test_fasta_crlf_and_blank_lines_inside_a_record(tmp_path):
    # fasta::Reader::read: `record.seq.push_str(self.line.trim_end())` for every line up to the next '>' — CR, blanks
    # and trailing spaces vanish, blank lines contribute nothing
    p = _write(tmp_path, "a.fa", b">r1 desc\r\nACGT  \r\n\r\nTTGA\r\n\n>r2\nGG\n\n\n")
    assert _as_list(*kio.read_fastx(p)) == [b"ACGTTTGA", b"GG"]


def test_fasta_first_line_must_be_a_header(tmp_path):
    # fasta::Reader::read: `if !self.line.starts_with('>') { Err("Expected > at record start.") }` — a blank FIRST line
    # is not skipped (the reference then panics on unwrap)
    from kmertools_b200 import KtbError
    with pytest.raises(KtbError, match="Expected > at record start"):
        kio.read_fastx(_write(tmp_path, "b.fa", b"\n>r1\nACGT\n"))


def test_fasta_empty_record_ends_the_iteration(tmp_path):
    # fasta::Records::next: `Ok(()) if record.is_empty() => None` — a bare ">" followed directly by another header
    # (no id, no description, no sequence) stops the iteration, whatever follows; with an id the empty record counts
    assert _as_list(*kio.read_fastx(_write(tmp_path, "c.fa", b">r1\nAC\n>\n>r3\nGG\n"))) == [b"AC"]
    assert _as_list(*kio.read_fastx(_write(tmp_path, "d.fa", b">r1\nAC\n>r2\n>r3\nGG\n"))) == [b"AC", b"", b"GG"]
    assert _as_list(*kio.read_fastx(_write(tmp_path, "e.fa", b">r1\nAC\n>\nTT\n>r3\nGG\n"))) == [b"AC", b"TT", b"GG"]


def test_fasta_trailing_header_with_every_batch_boundary(tmp_path):
    # a header as the LAST line (with and without a final newline) is a record with an empty sequence; batch limits
    # must not lose it (the batch loop used to stop on end-of-input while that header was still pending)
    for tail in (b">last\n", b">last"):
        p = _write(tmp_path, "f.fa", b">r1\nACGT\n>r2\nGGA\nTT\n" + tail)
        want = [b"ACGT", b"GGATT", b""]
        assert _as_list(*kio.read_fastx(p)) == want
        for max_records in (1, 2, 3):
            for batch_bytes in (4, 5, 7, 64):
                assert _as_list(*kio.read_fastx_batched(p, max_records, batch_bytes)) == want, (max_records, batch_bytes)


def test_batch_boundaries_do_not_change_the_records(tmp_path):
    rng = np.random.default_rng(11)
    seqs = [bytes(rng.choice(list(b"ACGTN"), size=int(n)).astype(np.uint8)) for n in rng.integers(0, 200, size=60)]
    fa = b"".join(b">s%d\n" % i + b"\n".join(s[j:j + 37] for j in range(0, len(s), 37)) + b"\n" for i, s in enumerate(seqs))
    fq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate(seqs) if s)
    pa, pq = _write(tmp_path, "g.fa", fa), _write(tmp_path, "g.fq", fq)
    for max_records, batch_bytes in ((1, 1 << 20), (7, 1 << 20), (1000, 64), (3, 150), (1000, 1)):
        assert _as_list(*kio.read_fastx_batched(pa, max_records, batch_bytes)) == seqs
        assert _as_list(*kio.read_fastx_batched(pq, max_records, batch_bytes)) == [s for s in seqs if s]


def test_fastq_multiline_and_marker_characters_in_quality(tmp_path):
    # fastq::Reader::read: sequence lines up to the first line starting with '+', then EXACTLY as many quality lines as
    # there were sequence lines — so a quality line may start with '@', '+' or '>'
    fq = b"@r1 d\nACGT\nTTGA\n+r1\n@III\n+>II\n@r2\nGG\n+\n>I\n"
    assert _as_list(*kio.read_fastx(_write(tmp_path, "h.fq", fq))) == [b"ACGTTTGA", b"GG"]
    # CRLF: trim_end on every line
    assert _as_list(*kio.read_fastx(_write(tmp_path, "i.fq", fq.replace(b"\n", b"\r\n")))) == [b"ACGTTTGA", b"GG"]


def test_fastq_errors_follow_the_reader(tmp_path):
    from kmertools_b200 import KtbError
    # Error::MissingAt for anything but '@' where a header is expected — including a blank line between records and a
    # blank line at the end of the file (read_line returns "\n", which is not empty)
    for data in (b"@r1\nAC\n+\nII\n\n@r2\nGG\n+\nII\n", b"@r1\nAC\n+\nII\n\n", b"r1\nAC\n+\nII\n"):
        with pytest.raises(KtbError, match="Expected @ at record start"):
            kio.read_fastx(_write(tmp_path, "j.fq", data))
    # Error::IncompleteRecord when the quality string is empty: end of input before '+', no sequence line, blank quality
    for data in (b"@r1\nACGT\n", b"@r1\n+\n", b"@r1\nAC\n+\n\n", b"@r1\nACGT"):
        with pytest.raises(KtbError, match="Incomplete record"):
            kio.read_fastx(_write(tmp_path, "k.fq", data))
    # a missing final newline after a complete record is fine
    assert _as_list(*kio.read_fastx(_write(tmp_path, "l.fq", b"@r1\nACGT\n+\nIIII"))) == [b"ACGT"]


def test_gzip_only_the_first_member_is_read(tmp_path):
    # flate2::read::GzDecoder decodes ONE member (MultiGzDecoder would continue); ktio/src/seq.rs:149 uses GzDecoder
    m1 = gzip.compress(b">r1\nACGT\n>r2\nGG\n")
    m2 = gzip.compress(b">r3\nTTTT\n")
    assert _as_list(*kio.read_fastx(_write(tmp_path, "m.fa.gz", m1 + m2))) == [b"ACGT", b"GG"]
    assert _as_list(*kio.read_fastx(_write(tmp_path, "n.fa.gz", m1))) == [b"ACGT", b"GG"]
    big = b"".join(b">s%d\n%s\n" % (i, b"ACGTTGCA" * 200) for i in range(3000))       # several inflate calls
    got = _as_list(*kio.read_fastx(_write(tmp_path, "o.fa.gz", gzip.compress(big) + m2)))
    assert len(got) == 3000 and all(s == b"ACGTTGCA" * 200 for s in got)
    from kmertools_b200 import KtbError
    with pytest.raises(KtbError):   # a truncated member is an error, as it is for GzDecoder
        kio.read_fastx(_write(tmp_path, "p.fa.gz", gzip.compress(big)[:-40]))
    with pytest.raises(KtbError):   # ".gz" that is not gzip
        kio.read_fastx(_write(tmp_path, "q.fa.gz", b">r1\nACGT\n"))


@pytest.mark.parametrize("mapped", [0, 1])
def test_span_writer_orders_blocks_and_handles_odd_sizes(tmp_path, mapped):
    """The drivers' output writer (csrc/span_writer.h): blocks submitted in order end up in order, whatever the block
    size, span boundaries (8 MB) and thread count; the file has exactly the submitted length."""
    import ctypes as C
    from kmertools_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(mapped)
    for nbytes, block, threads in ((0, 4096, 4), (1, 1, 1), (12345, 1000, 3), (9_000_001, 1_234_567, 8), (20_000_003, 17_000_000, 4)):
        data = rng.integers(0, 256, size=nbytes, dtype=np.uint8)
        p = tmp_path / f"w{mapped}_{nbytes}.bin"
        _lib.check(L.ktb_debug_span_write(str(p).encode(), data.ctypes.data if nbytes else None, nbytes, block, threads, mapped))
        assert p.read_bytes() == data.tobytes()
    # not a regular file: written sequentially whatever was asked for
    _lib.check(L.ktb_debug_span_write(b"/dev/null", data.ctypes.data, len(data), 4096, 4, mapped))
    from kmertools_b200 import KtbError
    with pytest.raises(KtbError, match="Unable to write"):
        _lib.check(L.ktb_debug_span_write(str(tmp_path / "no" / "dir" / "x").encode(), data.ctypes.data, 10, 4, 1, mapped))
