"""pykmertools.utils mirrors (pybindings/src/kmer.rs, kmer/src/lib.rs:19-50): pure integer helpers."""

_L = "ACGT"
_C = {"A": 0, "a": 0, "C": 1, "c": 1, "G": 2, "g": 2, "T": 3, "t": 3, "U": 3, "u": 3,
      "\x00": 0, "\x01": 1, "\x02": 2, "\x03": 3}


def to_acgt(kmer: int, k: int) -> str:
    """numeric_to_kmer, kmer/src/lib.rs:19-34."""
    return "".join(_L[(kmer >> (2 * (k - 1 - i))) & 3] for i in range(k))


def to_numeric(kmer: str) -> tuple[int, int]:
    """kmer_to_numeric, kmer/src/lib.rs:36-50: (forward, reverse-complement) codes."""
    k = len(kmer)
    if k > 32:   # pybindings/src/kmer.rs:57-63 (a u64 holds 32 bases)
        raise ValueError(f"Invalid k-mer length: {k}, must be <= 32")
    mask = (1 << (2 * k)) - 1
    f = r = 0
    for ch in kmer:
        c = _C.get(ch, 4)
        f = ((f << 2) | c) & mask
        r = (r >> 2) | ((c ^ 3) << (2 * (k - 1)))
    return f & 0xFFFFFFFFFFFFFFFF, r & 0xFFFFFFFFFFFFFFFF
