"""File-level front end: FASTA/FASTQ loading and the `kmertools comp oligo` driver (C ABI wrappers)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib


def read_fastx(path: str | os.PathLike, sniff: bool = False) -> tuple[np.ndarray, np.ndarray]:
    """(bases u8, offsets u64[n+1]) of a FASTA/FASTQ(.gz) file, parsed by the library's C++ reader
    (the feeder the CLI uses; restates ktio/src/seq.rs).  sniff=True decides FASTA/FASTQ from the first
    byte like the reference's batch driver, otherwise from the extension."""
    L = _lib.load()
    pb, po, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
    _lib.check(L.ktb_fastx_load(os.fsencode(str(path)), int(sniff), C.byref(pb), C.byref(po), C.byref(n)))
    try:
        offsets = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_uint64)), shape=(n.value + 1,)).copy()
        total = int(offsets[-1])
        if total:
            bases = np.ctypeslib.as_array(C.cast(pb, C.POINTER(C.c_uint8)), shape=(total,)).copy()
        else:
            bases = np.zeros(0, dtype=np.uint8)
    finally:
        L.ktb_free(pb)
        L.ktb_free(po)
    return bases, offsets


def read_fastx_batched(path: str | os.PathLike, max_records: int, batch_bytes: int, sniff: bool = False):
    """read_fastx through the drivers' batch loop with the given per-batch limits (test hook)."""
    L = _lib.load()
    pb, po, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
    _lib.check(L.ktb_debug_fastx_batches(os.fsencode(str(path)), int(sniff), int(max_records), int(batch_bytes),
                                         C.byref(pb), C.byref(po), C.byref(n)))
    try:
        offsets = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_uint64)), shape=(n.value + 1,)).copy()
        total = int(offsets[-1])
        bases = (np.ctypeslib.as_array(C.cast(pb, C.POINTER(C.c_uint8)), shape=(total,)).copy() if total
                 else np.zeros(0, dtype=np.uint8))
    finally:
        L.ktb_free(pb)
        L.ktb_free(po)
    return bases, offsets


def comp_oligo(in_path, out_path, k: int = 3, counts: bool = False, raw_count: bool = False,
               preset: str = "spc", header: bool = False, threads: int = 0, device: int = 0) -> dict:
    """`kmertools comp oligo -i in -o out [-c] [-k K] [-r] [-p preset] [-H] [-t N]` on the GPU
    (kmertools/src/args.rs:70-103,242-263).  Returns the driver's timing / volume statistics."""
    L = _lib.load()
    delim = {"csv": b",", "tsv": b"\t", "spc": b" "}[preset]
    o = _lib.FileOpts(os.fsencode(str(in_path)), os.fsencode(str(out_path)), int(k), int(not raw_count),
                      int(not counts), delim, int(header), int(threads), int(device))
    st = _lib.FileStats()
    _lib.check(L.ktb_comp_oligo_file(C.byref(o), C.byref(st)))
    return {f: getattr(st, f) for f, _ in st._fields_}


def comp_cgr(in_path, out_path, k: int, counts: bool = False, vec_size: int | None = None, threads: int = 0,
             device: int = 0) -> dict:
    """`kmertools comp cgr -i in -o out -k K [-c] [-v N]` (k-mer mode, kmertools/src/args.rs:105-128,264-283)."""
    L = _lib.load()
    if vec_size is None:
        vec_size = int((float(k) ** 4.0) ** 0.5)   # args.rs:268-271
    o = _lib.FileOpts(os.fsencode(str(in_path)), os.fsencode(str(out_path)), int(k), 1, int(not counts), b" ", 0,
                      int(threads), int(device))
    st = _lib.FileStats()
    _lib.check(L.ktb_comp_cgr_file(C.byref(o), int(vec_size), C.byref(st)))
    return {f: getattr(st, f) for f, _ in st._fields_}


def format6(q: float) -> str:
    """Host build of the GPU text formatter: Rust's format!("{:.6}", q) for q in [0, 1]."""
    buf = C.create_string_buffer(9)
    _lib.check(_lib.load().ktb_debug_format6(float(q), buf))
    return buf.raw[:8].decode()


def release_cached_buffers() -> None:
    """Free the pinned / device buffer sets comp_oligo and comp_cgr keep between calls."""
    _lib.load().ktb_release_cached_buffers()
