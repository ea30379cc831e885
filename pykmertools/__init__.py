"""Drop-in for the reference's `pykmertools` module (pip/src/lib.rs:31-40), oligo path only.

`import pykmertools as kt; kt.OligoComputer(4).vectorise_batch(seqs)` runs on the GPU through
libkmertools_b200.so.  Classes outside the oligo path (CgrComputer, MinimiserGenerator, ...) are not
provided: they are out of scope.
"""
from kmertools_b200.oligo import OligoComputer  # noqa: F401
from . import utils  # noqa: F401

__all__ = ["OligoComputer", "utils"]
