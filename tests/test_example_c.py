"""examples/vectorise.c: the header is usable from plain C, the library links, and the client behaves with and
without a GPU."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def build_example(tmp_path):
    from kmertools_b200 import build as kb
    kb.build()
    exe = tmp_path / "vectorise"
    cmd = ["/usr/bin/gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-o", str(exe),
           str(ROOT / "examples" / "vectorise.c"), f"-I{ROOT / 'include'}", f"-L{ROOT / 'kmertools_b200' / 'lib'}",
           "-lkmertools_b200", f"-Wl,-rpath,{ROOT / 'kmertools_b200' / 'lib'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_client_compiles_and_fails_loudly_without_gpu(tmp_path):
    from kmertools_b200 import _lib
    exe = build_example(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    if _lib.load().ktb_device_count() == 0:
        r = subprocess.run([str(exe), "4", "ACGTACGT"], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU path" in r.stderr and r.stdout == ""


@pytest.mark.gpu
def test_c_client_matches_oracle(tmp_path):
    from oracle import oracle as O
    exe = build_example(tmp_path)
    seqs = ["ACGTACGTNACGTTTGACCA", "GATTACA", "NNNN", "acgtuACGTU"]
    r = subprocess.run([str(exe), "4", *seqs], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [O.vectorise_one(s.encode(), 4, True, 1) for s in seqs]
    import numpy as np
    assert r.stdout.encode() == O.format_rows(np.stack(rows), True)
