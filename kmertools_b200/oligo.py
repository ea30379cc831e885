"""Host-side mirror of the reference's OligoComputer (pybindings/src/oligo.rs:7-99,
pybindings/pykmertools.pyi:173-235) on top of the C ABI.

Same constructor, method names, argument meaning, defaults and quirks:
  * OligoComputer(ksize)
  * vectorise_one(seq, norm=True, mins=True)   -> list[float]
  * vectorise_batch(seqs, norm=True, mins=True) -> list[list[float]]
  * get_header(mins=True)                       -> list[str]
  * raw mode (mins=False) normalises by 2 * #kmers, as pybindings/src/oligo.rs:58-62 does
  * never raises on sequence content (ambiguous bytes reset the k-mer window)
Supersets: vectorise_batch_array / vectorise_packed return numpy arrays without the per-float boxing,
vectorise_device works on CUDA tensors in place.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import NORM_CLI, NORM_COUNTS, NORM_PY, OUT_F32, OUT_F64, OUT_U32

_DTYPES = {OUT_U32: np.uint32, OUT_F32: np.float32, OUT_F64: np.float64}
_CODES = {np.dtype(np.uint32): OUT_U32, np.dtype(np.float32): OUT_F32, np.dtype(np.float64): OUT_F64}


class HostBuffer:
    """Page-locked host memory from ktb_host_alloc / ktb_host_alloc_near, viewed as a numpy array."""

    def __init__(self, shape, dtype, near_device: int | None = None):
        """near_device: bind the pages to the NUMA node of that GPU (ktb_host_alloc_near) instead of wherever the
        calling thread happens to run — what a process that feeds one of several GPUs wants."""
        self._lib = _lib.load()
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        if near_device is None:
            self.ptr = self._lib.ktb_host_alloc(max(nbytes, 1))
        else:
            self.ptr = self._lib.ktb_host_alloc_near(max(nbytes, 1), int(near_device))
        if not self.ptr:
            _lib.check(_lib.KTB_ERR_NOMEM)
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape, dtype=np.int64))
                                   ).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.ktb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _pack(seqs: Sequence) -> tuple[np.ndarray, np.ndarray]:
    """list of str/bytes -> (bases u8, offsets u64[n+1]); str is encoded as UTF-8 like Rust's as_bytes()."""
    bs = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        np.cumsum([len(b) for b in bs], out=offsets[1:])
    joined = b"".join(bs)
    bases = np.frombuffer(joined, dtype=np.uint8) if joined else np.zeros(0, dtype=np.uint8)
    return bases, offsets


class OligoComputer:
    """Computer for generating oligonucleotide frequency vectors (GPU)."""

    def __init__(self, ksize: int, device: int = 0):
        self._lib = _lib.load()
        self.ksize = int(ksize)
        h = C.c_void_p()
        _lib.check(self._lib.ktb_oligo_create(self.ksize, int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ktb_oligo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference interface
    def vectorise_one(self, seq: str, norm: bool = True, mins: bool = True) -> list[float]:
        return self.vectorise_batch([seq], norm, mins)[0]

    def vectorise_batch(self, seqs: Iterable[str], norm: bool = True, mins: bool = True) -> list[list[float]]:
        return self.vectorise_batch_array(list(seqs), norm, mins, dtype=np.float64).tolist()

    def get_header(self, mins: bool = True) -> list[str]:
        d = self.dim(mins)
        buf = C.create_string_buffer(d * self.ksize + 1)
        _lib.check(self._lib.ktb_oligo_header(self._h, int(mins), buf, d * self.ksize))
        raw = buf.raw[: d * self.ksize].decode()
        return [raw[i * self.ksize:(i + 1) * self.ksize] for i in range(d)]

    # ------------------------------------------------------------------ supersets
    def dim(self, mins: bool = True) -> int:
        return int(self._lib.ktb_oligo_dim(self._h, int(mins)))

    def kmer_pos_maps(self):
        """(pos_map[4^k], pos_to_kmer[count], count) — KmerGenerator::kmer_pos_maps, kmer.rs:54-73."""
        n = 4 ** self.ksize
        pm = np.zeros(n, dtype=np.uint64)
        pk = np.zeros(self.dim(True), dtype=np.uint64)
        cnt = C.c_uint64()
        _lib.check(self._lib.ktb_oligo_pos_maps(self._h, pm.ctypes.data, pk.ctypes.data, C.byref(cnt)))
        return pm, pk, int(cnt.value)

    def set_option(self, key: str, value: int) -> None:
        _lib.check(self._lib.ktb_oligo_set_option(self._h, key.encode(), int(value)))

    def stats(self) -> dict:
        st = _lib.Stats()
        _lib.check(self._lib.ktb_oligo_last_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def vectorise_batch_array(self, seqs: Sequence, norm: bool = True, mins: bool = True,
                              dtype=np.float64) -> np.ndarray:
        bases, offsets = _pack(seqs)
        return self.vectorise_packed(bases, offsets, norm_mode=NORM_PY if norm else NORM_COUNTS, mins=mins,
                                     dtype=dtype)

    def vectorise_packed(self, bases: np.ndarray, offsets: np.ndarray, norm_mode: int = NORM_CLI,
                         mins: bool = True, dtype=np.float32, out: np.ndarray | None = None,
                         totals: np.ndarray | None = None) -> np.ndarray:
        """Rows for sequences bases[offsets[i]:offsets[i+1]] (host buffers, chunked H2D/compute/D2H)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        d = self.dim(mins)
        code = _CODES[np.dtype(dtype)]
        if out is None:
            out = np.empty((n, d), dtype=_DTYPES[code])
        assert out.shape == (n, d) and out.dtype == _DTYPES[code] and out.flags.c_contiguous
        tptr = None
        if totals is not None:
            assert totals.dtype == np.uint64 and totals.shape == (n,) and totals.flags.c_contiguous
            tptr = totals.ctypes.data
        bptr = bases.ctypes.data if bases.size else None
        _lib.check(self._lib.ktb_oligo_vectorise(self._h, bptr, offsets.ctypes.data, n, int(mins), int(norm_mode),
                                                 code, out.ctypes.data if n else None, tptr))
        return out

    def vectorise_tensors(self, bases, offsets, norm_mode: int = NORM_CLI, mins: bool = True, dtype=None,
                          out=None, totals=None):
        """Zero-copy surface for torch users ("next" row N4): `bases` (uint8) and `offsets` (int64, n+1) are
        CUDA tensors on this handle's GPU, the result is a CUDA tensor (n x dim) — nothing crosses PCIe.
        The work is enqueued on torch's current stream."""
        import torch
        assert bases.is_cuda and offsets.is_cuda and bases.dtype == torch.uint8 and offsets.dtype == torch.int64
        assert bases.device.index == self.device and bases.is_contiguous() and offsets.is_contiguous()
        dtype = dtype or torch.float32
        code = {torch.int32: OUT_U32, torch.float32: OUT_F32, torch.float64: OUT_F64}[dtype]
        n = offsets.numel() - 1
        if out is None:
            out = torch.empty((n, self.dim(mins)), dtype=dtype, device=bases.device)
        assert out.is_cuda and out.dtype == dtype and out.is_contiguous() and tuple(out.shape) == (n, self.dim(mins))
        total = int(offsets[-1]) if n >= 0 and offsets.numel() else 0
        tptr = None
        if totals is not None:
            assert totals.is_cuda and totals.dtype == torch.int64 and totals.numel() == n
            tptr = totals.data_ptr()
        stream = torch.cuda.current_stream(bases.device).cuda_stream
        self.vectorise_device(bases.data_ptr() if total else 0, offsets.data_ptr(), n, total, out.data_ptr() if n else 0,
                              norm_mode=norm_mode, mins=mins, out_dtype=code, d_totals=tptr, stream=stream)
        return out

    def vectorise_batch_tensor(self, seqs: Sequence, norm: bool = True, mins: bool = True, dtype=None):
        """vectorise_batch with the rows left on the GPU as a torch tensor (reference semantics, PY norm)."""
        import torch
        bases, offsets = _pack(seqs)
        dev = torch.device("cuda", self.device)
        tb = torch.from_numpy(bases.copy() if bases.size else np.zeros(16, np.uint8)).to(dev)
        to = torch.from_numpy(offsets.astype(np.int64)).to(dev)
        return self.vectorise_tensors(tb, to, NORM_PY if norm else NORM_COUNTS, mins, dtype)

    def vectorise_device(self, d_bases: int, d_offsets: int, n: int, total_bases: int, d_out: int,
                         norm_mode: int = NORM_CLI, mins: bool = True, out_dtype: int = OUT_F32,
                         d_totals: int | None = None, stream: int | None = None) -> None:
        """Device-pointer entry point (raw addresses, e.g. torch.Tensor.data_ptr()); asynchronous."""
        _lib.check(self._lib.ktb_oligo_vectorise_device(self._h, d_bases, d_offsets, n, total_bases, int(mins),
                                                        int(norm_mode), int(out_dtype), d_out, d_totals, stream))


class MultiOligoComputer:
    """All GPUs of the box behind one call (ktb_multi_*, csrc/multi.cu): the batch is cut into one contiguous range
    of sequences per device, balanced by bases; every device writes its own slab of rows; row order = input order,
    like the reference's rayon fan-out (pybindings/src/oligo.rs:77-81).  devices=None takes every visible GPU."""

    def __init__(self, ksize: int, devices: Sequence[int] | None = None):
        self._lib = _lib.load()
        self.ksize = int(ksize)
        devs = list(devices) if devices is not None else []
        arr = (C.c_int * max(1, len(devs)))(*devs)
        m = C.c_void_p()
        _lib.check(self._lib.ktb_multi_create(self.ksize, arr, len(devs), C.byref(m)))
        self._m = m
        self.ndev = int(self._lib.ktb_multi_device_count(m))

    def close(self):
        if getattr(self, "_m", None):
            self._lib.ktb_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dim(self, mins: bool = True) -> int:
        return int(self._lib.ktb_oligo_dim(self._lib.ktb_multi_handle(self._m, 0), int(mins)))

    def set_option(self, key: str, value: int) -> None:
        for i in range(self.ndev):
            _lib.check(self._lib.ktb_oligo_set_option(self._lib.ktb_multi_handle(self._m, i), key.encode(), int(value)))

    def alloc_rows(self, offsets: np.ndarray, mins: bool = True, dtype=np.float32) -> "HostRows":
        """Page-locked (n, dim) output whose slab of device i lives on that device's NUMA node."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        return HostRows(self, offsets, mins, dtype)

    def vectorise_packed(self, bases: np.ndarray, offsets: np.ndarray, norm_mode: int = NORM_CLI, mins: bool = True,
                         dtype=np.float32, out: np.ndarray | None = None, totals: np.ndarray | None = None) -> np.ndarray:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        d = self.dim(mins)
        code = _CODES[np.dtype(dtype)]
        if out is None:
            out = np.empty((n, d), dtype=_DTYPES[code])
        assert out.shape == (n, d) and out.dtype == _DTYPES[code] and out.flags.c_contiguous
        tptr = None
        if totals is not None:
            assert totals.dtype == np.uint64 and totals.shape == (n,) and totals.flags.c_contiguous
            tptr = totals.ctypes.data
        _lib.check(self._lib.ktb_multi_vectorise(self._m, bases.ctypes.data if bases.size else None, offsets.ctypes.data, n,
                                                 int(mins), int(norm_mode), code, out.ctypes.data if n else None, tptr))
        return out

    def vectorise_batch(self, seqs: Iterable[str], norm: bool = True, mins: bool = True) -> list[list[float]]:
        """Reference semantics (pybindings/src/oligo.rs:77-81) over all devices."""
        bases, offsets = _pack(list(seqs))
        return self.vectorise_packed(bases, offsets, NORM_PY if norm else NORM_COUNTS, mins, np.float64).tolist()

    def device_stats(self, i: int) -> dict:
        st = _lib.Stats()
        lo, hi = C.c_uint64(), C.c_uint64()
        _lib.check(self._lib.ktb_multi_last_stats(self._m, int(i), C.byref(st), C.byref(lo), C.byref(hi)))
        d = {f: getattr(st, f) for f, _ in st._fields_}
        d.update(first_row=int(lo.value), end_row=int(hi.value))
        return d


class HostRows:
    """Output buffer from ktb_multi_alloc_rows, viewed as a numpy array."""

    def __init__(self, multi: MultiOligoComputer, offsets: np.ndarray, mins: bool, dtype):
        self._lib = multi._lib
        n = len(offsets) - 1
        code = _CODES[np.dtype(dtype)]
        d = multi.dim(mins)
        self.ptr = self._lib.ktb_multi_alloc_rows(multi._m, offsets.ctypes.data, n, int(mins), code)
        if not self.ptr:
            _lib.check(_lib.KTB_ERR_NOMEM)
        nbytes = max(1, n * d * np.dtype(dtype).itemsize)
        buf = (C.c_uint8 * nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=_DTYPES[code], count=n * d).reshape(n, d)

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.ktb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def shard_bounds(offsets: np.ndarray, parts: int) -> np.ndarray:
    """bounds[0..parts] of the library's partition (ktb_shard_bounds): part r = sequences [bounds[r], bounds[r+1])."""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    out = np.zeros(int(parts) + 1, dtype=np.uint64)
    _lib.check(_lib.load().ktb_shard_bounds(offsets.ctypes.data, len(offsets) - 1, int(parts), out.ctypes.data))
    return out
