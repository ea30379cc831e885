"""KmerGenerator on the GPU (ktb_kmer_pairs) against the reference's KATs and the CPU oracle.  Needs a B200: -m gpu."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import random_batch

pytestmark = pytest.mark.gpu

from kmertools_b200 import KmerGenerator, kmer_pairs  # noqa: E402


def pairs(seq, k):
    f, r = kmer_pairs(seq, k)
    return list(zip(f.tolist(), r.tolist()))


def test_reference_kats():
    """kmer/src/kmer.rs:113-145."""
    assert pairs(b"ACGT", 2) == [(1, 11), (6, 6), (11, 1)]
    assert pairs(b"ACNGTT", 2) == [(1, 11), (11, 1), (15, 0)]
    assert list(KmerGenerator("ACGT", 2)) == [(1, 11), (6, 6), (11, 1)]


def test_reference_python_test_unmodified():
    """tests/test_kmers.py of the reference, against the pykmertools drop-in."""
    import pykmertools as kt
    from pykmertools import utils as ktutils
    kmer_gen = kt.KmerGenerator("ACGTCC", 3)
    kmers = list(kmer_gen)
    kmers_acgt = ["ACG", "CGT", "GTC", "TCC"]
    assert len(kmers) == len(kmers_acgt)
    for (fmer, _), acgt_mer in zip(kmers, kmers_acgt):
        assert ktutils.to_acgt(fmer, len(acgt_mer)) == acgt_mer


def test_pos_maps_match_reference_test():
    """kmer/src/kmer.rs:156-176 (pos_map_test) through the Python mirror."""
    pos_map, pos_to_kmer, count = KmerGenerator("", 4).kmer_pos_maps()
    assert count == 136 and len(pos_to_kmer) == 136 and len(pos_map) == 256
    assert sum(1 for p in pos_map if p > 0) == 135 and max(pos_map) < 136
    assert pos_map[0] == 0 and pos_map[0b11111111] == 0 and pos_map[0b11] == 0b11
    om, ok, oc = O.kmer_pos_maps(4)
    assert pos_map == om.tolist() and [pos_to_kmer[j] for j in range(count)] == ok.tolist() and oc == count


@pytest.mark.parametrize("k", [1, 2, 3, 5, 15, 16, 17, 31])
def test_pairs_match_oracle(k):
    rng = np.random.default_rng(1000 + k)
    lengths = [0, 1, k - 1, k, k + 1, 15, 16, 17, 4095, 4096, 4097, 4096 + k, 70000]
    for n in lengths:
        bases, _ = random_batch(rng, [n], noise=0.01, n_runs=0.5)
        seq = bases[:n].tobytes()
        assert pairs(seq, k) == O.kmers(seq, k), (k, n)


def test_alphabet_and_raw_codes():
    """Lower case, U, raw 0..3 bytes are bases; everything else resets the window (kmer/src/kmer.rs:6-15)."""
    for seq in (b"acgu", b"ACGT", bytes([0, 1, 2, 3])):
        assert pairs(seq, 4) == [(27, 27)]
    seq = bytes(range(256)) * 3
    for k in (1, 2, 4):
        assert pairs(seq, k) == O.kmers(seq, k)
    assert pairs(b"ACGTRACGT", 4) == O.kmers(b"ACGTRACGT", 4) and len(pairs(b"ACGTRACGT", 4)) == 2
    assert pairs(b"NNNNNNNN", 3) == [] and pairs(b"", 3) == [] and pairs(b"AC", 3) == []


def test_capacity_smaller_than_count():
    """cap limits what is written, *count still reports every valid window."""
    import ctypes as C
    from kmertools_b200 import _lib
    L = _lib.load()
    seq = np.frombuffer(b"ACGTACGTACGTACGT", dtype=np.uint8)
    f = np.zeros(5, dtype=np.uint64)
    r = np.zeros(5, dtype=np.uint64)
    n = C.c_uint64()
    _lib.check(L.ktb_kmer_pairs(seq.ctypes.data, seq.size, 3, 0, f.ctypes.data, r.ctypes.data, 5, C.byref(n)))
    want = O.kmers(seq.tobytes(), 3)
    assert n.value == len(want) == 14
    assert list(zip(f.tolist(), r.tolist())) == want[:5]
    _lib.check(L.ktb_kmer_pairs(seq.ctypes.data, seq.size, 3, 0, None, None, 0, C.byref(n)))
    assert n.value == 14


def test_device_entry_point():
    import ctypes as C
    import torch
    from kmertools_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(5)
    bases, _ = random_batch(rng, [300000], noise=0.001, n_runs=0.2)
    seq = bases[:300000]
    d_seq = torch.from_numpy(seq.copy()).cuda()
    k = 21
    cap = seq.size - k + 1
    d_f = torch.empty(cap, dtype=torch.int64, device="cuda")
    d_r = torch.empty(cap, dtype=torch.int64, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.ktb_kmer_pairs_device(d_seq.data_ptr(), seq.size, k, d_f.data_ptr(), d_r.data_ptr(), cap,
                                       d_n.data_ptr(), st))
    torch.cuda.synchronize()
    m = int(d_n.item())
    want = O.kmers(seq.tobytes(), k)
    assert m == len(want)
    got = list(zip(d_f[:m].cpu().numpy().astype(np.uint64).tolist(), d_r[:m].cpu().numpy().astype(np.uint64).tolist()))
    assert got == want
