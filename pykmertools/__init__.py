"""Drop-in for the reference's `pykmertools` module (pip/src/lib.rs:31-40), k-mer / oligo path only.

`import pykmertools as kt; kt.OligoComputer(4).vectorise_batch(seqs)` and `kt.KmerGenerator(seq, k)` run on the
GPU through libkmertools_b200.so.  Classes outside that path (CgrComputer, MinimiserGenerator, ...) are not
provided: they are out of scope.
"""
from kmertools_b200.oligo import OligoComputer  # noqa: F401
from kmertools_b200.kmers import KmerGenerator  # noqa: F401
from . import utils  # noqa: F401

__all__ = ["OligoComputer", "KmerGenerator", "utils"]
