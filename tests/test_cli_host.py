"""CPU-side checks of the CLI boundary and of the ABI struct layouts (no GPU needed)."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest

from kmertools_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "kmertools_b200" / "bin" / "kmertools"


def run(*args):
    return subprocess.run([str(BIN), *map(str, args)], capture_output=True, text=True)


def test_cli_argument_errors(golden, tmp_path):
    """clap-like behaviour of kmertools/src/args.rs:70-103: unknown subcommand / missing / out-of-range -> exit 2."""
    assert BIN.exists(), "build the CLI first (python -m kmertools_b200.build)"
    assert run().returncode == 2
    assert run("ctr", "-i", "x").returncode == 2
    r = run("comp", "oligo", "-i", golden / "reads.fa")
    assert r.returncode == 2 and "--output" in r.stderr
    r = run("comp", "oligo", "-i", golden / "reads.fa", "-o", tmp_path / "o", "-k", "8")
    assert r.returncode == 2 and "3..=7" in r.stderr
    r = run("comp", "oligo", "-i", golden / "reads.fa", "-o", tmp_path / "o", "-p", "xml")
    assert r.returncode == 2 and "csv, tsv, spc" in r.stderr
    assert run("comp", "oligo", "--help").returncode == 0
    r = run("comp", "cgr", "-i", golden / "reads.fa", "-o", tmp_path / "o")
    assert r.returncode == 2 and "whole-sequence CGR" in r.stderr


def test_cli_reports_missing_gpu_like_a_reference_error(golden, tmp_path):
    """Without a GPU the driver fails loudly (no CPU fallback); the CLI prints `Error: ...` and exits 0 like
    the reference does for its own errors (args.rs:260-262)."""
    if _lib.load().ktb_device_count() > 0:
        pytest.skip("a GPU is present")
    r = run("comp", "oligo", "-i", golden / "reads.fa", "-o", tmp_path / "o", "-k", "4")
    assert r.returncode == 0 and r.stderr.startswith("Error:") and "no CUDA device" in r.stderr


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of the ABI structs as the C compiler sees them == the ctypes mirrors."""
    src = tmp_path / "probe.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "kmertools_b200.h"
int main(void) {
    printf("%zu %zu %zu\\n", sizeof(ktb_stats), sizeof(ktb_file_opts), sizeof(ktb_file_stats));
    printf("%zu %zu %zu %zu\\n", offsetof(ktb_file_opts, k), offsetof(ktb_file_opts, delim),
           offsetof(ktb_file_opts, header), offsetof(ktb_file_opts, device));
    printf("%zu %zu\\n", offsetof(ktb_stats, launches), offsetof(ktb_file_stats, launches));
    return 0;
}''')
    exe = tmp_path / "probe"
    subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = list(map(int, out))
    want = [C.sizeof(_lib.Stats), C.sizeof(_lib.FileOpts), C.sizeof(_lib.FileStats),
            _lib.FileOpts.k.offset, _lib.FileOpts.delim.offset, _lib.FileOpts.header.offset,
            _lib.FileOpts.device.offset, _lib.Stats.launches.offset, _lib.FileStats.launches.offset]
    assert got == want
