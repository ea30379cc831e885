"""KmerGenerator mirror (pybindings/src/kmer.rs:7-44) on top of ktb_kmer_pairs.

The reference iterates `KmerGenerator::next` (kmer/src/kmer.rs:80-106) one window at a time; here the whole
sequence is enumerated by one GPU call on first use and the iterator walks the result.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def kmer_pairs(seq, ksize: int, device: int = 0) -> tuple[np.ndarray, np.ndarray]:
    """(forward, reverse-complement) codes of every valid window of `seq`, in position order, as two uint64
    arrays.  `seq` is str / bytes / a uint8 array (ASCII or raw 0..3 codes, kmer/src/kmer.rs:6-15)."""
    if isinstance(seq, str):
        seq = seq.encode("utf-8")   # Rust's String::as_bytes (pybindings/src/kmer.rs:22-29), as OligoComputer does
    buf = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray, memoryview)) \
        else np.ascontiguousarray(seq, dtype=np.uint8)
    if not 1 <= int(ksize) <= 31:
        raise ValueError(f"ksize must be in 1..31 (got {ksize})")
    lib = _lib.load()
    cap = max(0, buf.size - int(ksize) + 1)
    f = np.empty(cap, dtype=np.uint64)
    r = np.empty(cap, dtype=np.uint64)
    count = C.c_uint64(0)
    _lib.check(lib.ktb_kmer_pairs(buf.ctypes.data if buf.size else None, buf.size, int(ksize), int(device),
                                  f.ctypes.data if cap else None, r.ctypes.data if cap else None, cap,
                                  C.byref(count)))
    return f[:count.value], r[:count.value]


class KmerGenerator:
    """An iterator object to generate k-mers as (forward, reverse) numeric kmer tuples
    (pybindings/pykmertools.pyi:98-134)."""

    def __init__(self, seq: str, ksize: int, device: int = 0) -> None:
        if not 1 <= int(ksize) <= 31:
            raise ValueError(f"ksize must be in 1..31 (got {ksize})")
        self._seq = seq
        self.ksize = int(ksize)
        self._device = device
        self._pairs = None
        self._i = 0

    def __iter__(self):
        return self

    def __next__(self) -> tuple[int, int]:
        if self._pairs is None:
            self._pairs = kmer_pairs(self._seq, self.ksize, self._device)
        f, r = self._pairs
        if self._i >= len(f):
            raise StopIteration
        i = self._i
        self._i += 1
        return int(f[i]), int(r[i])

    def kmer_pos_maps(self) -> tuple[list[int], dict[int, int], int]:
        """KmerGenerator::kmer_pos_maps (kmer/src/kmer.rs:54-73): (canonical code -> column, column -> canonical
        code, number of columns)."""
        from .oligo import OligoComputer
        if self.ksize > 12:
            raise ValueError("kmer_pos_maps needs a dense table of 4^ksize entries: ksize must be <= 12")
        pos_map, pos_to_kmer, count = OligoComputer(self.ksize, device=self._device).kmer_pos_maps()
        return pos_map.tolist(), {j: int(c) for j, c in enumerate(pos_to_kmer)}, count
