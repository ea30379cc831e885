"""`kmertools comp oligo` drop-in (C++ CLI over the C ABI) against the reference's golden files and the
oracle.  Restates composition/src/oligo.rs:311-432 (vec_mmap_test, vec_batch_*_test, *_with_header_test)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "kmertools_b200" / "bin" / "kmertools"


def run(args, stdin=None):
    return subprocess.run([str(BIN), "comp", "oligo", *map(str, args)], input=stdin, capture_output=True)


@pytest.mark.parametrize("fname", ["reads.fa", "reads.fq", "reads.fq.gz"])
def test_config1_golden_norm(golden, tmp_path, fname):
    """BASELINE config 1: kmertools comp oligo k=4 canonical normalised on the repo fixture."""
    out = tmp_path / "out.kmers"
    r = run(["-i", golden / fname, "-o", out, "-k", 4])
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    assert out.read_bytes() == (golden / "expected_fa.kmers").read_bytes()


def test_golden_counts_header_and_presets(golden, tmp_path):
    out = tmp_path / "o"
    assert run(["-i", golden / "reads.fa", "-o", out, "-k", 4, "-c"]).returncode == 0
    assert out.read_bytes() == (golden / "expected_fa_batch_unnorm.kmers").read_bytes()
    assert run(["-i", golden / "reads.fa", "-o", out, "-k", 4, "-H"]).returncode == 0
    assert out.read_bytes() == (golden / "expected_fa_header.kmers").read_bytes()
    spc = (golden / "expected_fa.kmers").read_bytes()
    run(["-i", golden / "reads.fa", "-o", out, "-k", 4, "-p", "csv"])
    assert out.read_bytes() == spc.replace(b" ", b",")
    run(["-i", golden / "reads.fa", "-o", out, "-k", 4, "--preset", "tsv", "-t", 8])
    assert out.read_bytes() == spc.replace(b" ", b"\t")


def test_stdin_input(golden, tmp_path):
    out = tmp_path / "o"
    r = run(["-i", "-", "-o", out, "-k", 4], stdin=(golden / "reads.fq").read_bytes())
    assert r.returncode == 0
    assert out.read_bytes() == (golden / "expected_fa.kmers").read_bytes()


def test_cli_errors_follow_the_reference(golden, tmp_path):
    r = run(["-i", tmp_path / "nope.fa", "-o", tmp_path / "o", "-k", 4])
    assert r.returncode == 0 and b"Error: Unable to open" in r.stderr     # args.rs:260-262 prints, exits 0
    assert run(["-i", golden / "reads.fa", "-o", tmp_path / "o", "-k", 9]).returncode == 2   # clap range 3..=7
    assert run(["-i", golden / "reads.fa"]).returncode == 2


@pytest.mark.parametrize("k", [3, 4, 5, 6, 7])
def test_random_file_matches_oracle_text(tmp_path, k):
    rng = np.random.default_rng(k)
    lengths = np.r_[rng.integers(0, 400, size=400), [30000, 0, 3, 150, 150]]
    fa = tmp_path / "r.fa"
    with open(fa, "wb") as fh:
        for i, n in enumerate(lengths):
            s = bytes(rng.choice(list(b"ACGTNacgtR"), size=int(n), p=[.22, .22, .22, .22, .02, .02, .02, .02, .02, .02]).astype(np.uint8))
            fh.write(b">s%d\n" % i)
            for j in range(0, len(s), 70):
                fh.write(s[j:j + 70] + b"\n")
    out = tmp_path / "o"
    for flags, kw in ((["-k", k], {}), (["-k", k, "-c"], {"norm": False}), (["-k", k, "-r"], {"canonical": False}),
                      (["-k", k, "-r", "-c", "-H", "-p", "csv"], {"canonical": False, "norm": False, "with_header": True, "delim": ","})):
        r = run(["-i", fa, "-o", out, *flags])
        assert r.returncode == 0 and r.stderr == b"", r.stderr
        assert out.read_bytes() == O.comp_oligo_text(fa, k, **kw), flags


def test_python_driver_and_many_batches(tmp_path):
    """comp_oligo() through ctypes; enough records for several GPU batches (rows stay in input order)."""
    from kmertools_b200 import io as kio
    rng = np.random.default_rng(2)
    fq = tmp_path / "big.fq"
    n = 12_000
    with open(fq, "wb") as fh:
        arr = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, 100))]
        for i in range(n):
            fh.write(b"@r\n" + arr[i].tobytes() + b"\n+\n" + b"I" * 100 + b"\n")
    out = tmp_path / "o"
    st = kio.comp_oligo(fq, out, k=7)          # 8192 cols * 9 B = 73.7 kB/row -> ~3600 rows per batch
    assert st["records"] == n and st["bases"] == n * 100 and st["launches"] >= 8
    data = out.read_bytes()
    assert len(data) == n * 8192 * 9
    rows, _ = O.vectorise_batch(arr.reshape(-1).copy(), np.arange(n + 1, dtype=np.uint64) * 100, 7, True, 1)
    for i in (0, 1, 3639, 3640, 3641, 3642, n // 2, n - 1):
        assert data[i * 73728:(i + 1) * 73728] == O.format_rows(rows[i:i + 1], True), i


def test_comp_cgr_kmer_mode_golden(golden, tmp_path):
    """oligo_cgr_complete_unnorm_test (composition/src/oligocgr.rs:219-238): reads.fq, k=4, vecsize 16, counts."""
    out = tmp_path / "o.cgr"
    r = subprocess.run([str(BIN), "comp", "cgr", "-i", str(golden / "reads.fq"), "-o", str(out), "-k", "4", "-c"],
                       capture_output=True)
    assert r.returncode == 0 and r.stderr == b"", r.stderr
    assert out.read_bytes() == (golden / "expected_reads.k4.cgr").read_bytes()


def test_comp_cgr_normalised_values(golden, tmp_path):
    """Normalised k-mer CGR: freq is Rust's `{}` of the f64 quotient = Python's repr (shortest round-trip)."""
    from kmertools_b200 import io as kio
    out = tmp_path / "o.cgr"
    kio.comp_cgr(golden / "reads.fa", out, k=4)
    seqs = [s for _, s in O.read_fastx(golden / "reads.fa")]
    rows, _ = O.vectorise_batch(*O.pack(seqs), 4, True, 1)
    lines = out.read_text().splitlines()
    assert len(lines) == 2
    for ln, row in zip(lines, rows):
        trip = [t.strip("()").split(",") for t in ln.split(" ")]
        assert len(trip) == 136
        assert trip[0][:2] == ["0.5", "0.5"]                 # AAAA -> (0.5, 0.5), oligocgr.rs:205-206
        for (x, y, f), want in zip(trip, row):
            w = repr(float(want))
            assert f == (w[:-2] if w.endswith(".0") else w), (f, w)
