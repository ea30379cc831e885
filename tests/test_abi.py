"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/kmertools_b200.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from kmertools_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "kmertools_b200.h").read_text()
    declared = set(re.findall(r"\b(ktb_[a-z0-9_]+)\s*\(", header))
    declared -= {"ktb_oligo", "ktb_stats"}
    assert declared, "no declarations parsed"
    L = _lib.load()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert L.ktb_abi_version() == 1


def test_library_is_standalone():
    """cudart is linked statically: the only CUDA dependency is the driver (loaded lazily)."""
    import subprocess
    out = subprocess.run(["ldd", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "libcudart" not in out and "libtorch" not in out and "oracle" not in out


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure: no product source may import, link or even mention it."""
    files = list((ROOT / "kmertools_b200").rglob("*.py")) + list((ROOT / "pykmertools").rglob("*.py")) + \
        [p for p in (ROOT / "kmertools_b200" / "csrc").glob("*") if p.is_file()]
    assert files
    for p in files:
        assert "oracle" not in p.read_text(errors="ignore").lower(), p


def test_argument_errors_without_gpu():
    L = _lib.load()
    h = C.c_void_p()
    assert L.ktb_oligo_create(0, 0, C.byref(h)) == _lib.KTB_ERR_ARG
    assert L.ktb_oligo_create(13, 0, C.byref(h)) == _lib.KTB_ERR_ARG
    assert b"k must be" in L.ktb_last_error()
    if L.ktb_device_count() == 0:
        assert L.ktb_oligo_create(4, 0, C.byref(h)) == _lib.KTB_ERR_NODEVICE
        assert b"no CPU path" in L.ktb_last_error()
        from kmertools_b200 import OligoComputer, KtbError
        with pytest.raises(KtbError):
            OligoComputer(4)


def test_kmer_pairs_argument_errors_without_gpu():
    L = _lib.load()
    n = C.c_uint64(7)
    seq = (C.c_uint8 * 4)(*b"ACGT")
    assert L.ktb_kmer_pairs(seq, 4, 0, 0, None, None, 0, C.byref(n)) == _lib.KTB_ERR_ARG
    assert L.ktb_kmer_pairs(seq, 4, 32, 0, None, None, 0, C.byref(n)) == _lib.KTB_ERR_ARG
    assert b"1..31" in L.ktb_last_error()
    assert L.ktb_kmer_pairs(seq, 4, 2, 0, None, None, 0, None) == _lib.KTB_ERR_ARG
    assert L.ktb_kmer_pairs(seq, 4, 2, 0, None, None, 3, C.byref(n)) == _lib.KTB_ERR_ARG
    if L.ktb_device_count() == 0:
        assert L.ktb_kmer_pairs(seq, 4, 2, 0, None, None, 0, C.byref(n)) == _lib.KTB_ERR_NODEVICE
        from kmertools_b200 import KmerGenerator, KtbError
        with pytest.raises(KtbError):
            next(KmerGenerator("ACGT", 2))
    with pytest.raises(ValueError):
        from kmertools_b200 import KmerGenerator
        KmerGenerator("ACGT", 0)


def test_utils_mirror():  # tests/test_utils.py of the reference
    from pykmertools import utils
    assert utils.to_acgt(111, 5) == "ACGTT"
    assert utils.to_numeric("ACGTT") == (111, 27)
    assert utils.to_numeric("T" * 32) == (2 ** 64 - 1, 0)
    with pytest.raises(ValueError, match="Invalid k-mer length: 33, must be <= 32"):   # pybindings/src/kmer.rs:57-63
        utils.to_numeric("A" * 33)


def test_documented_options_are_the_accepted_options():
    """Every option named in the header's ktb_oligo_set_option comment is handled in csrc/api.cu and vice versa, so the
    documentation of the tuning knobs cannot drift from the code."""
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    header = (root / "include" / "kmertools_b200.h").read_text()
    block = header[:header.index("int ktb_oligo_set_option(")]
    block = block[block.rindex("/*"):]
    documented = set(re.findall(r'^ \*\s+"([a-z0-9_]+)"', block, flags=re.M)) | set(re.findall(r'and "([a-z0-9_]+)"', block))
    api = (root / "kmertools_b200" / "csrc" / "api.cu").read_text()
    body = api[api.index("int ktb_oligo_set_option("):]
    body = body[:body.index("\n}\n")]
    accepted = set(re.findall(r'strcmp\(key, "([a-z0-9_]+)"\)', body))
    assert accepted, "no options found in api.cu"
    assert documented == accepted, f"undocumented: {sorted(accepted - documented)}; unknown to the code: {sorted(documented - accepted)}"
