#!/usr/bin/env python
"""Phase timing of the file-level driver on a synthetic FASTQ (runs on the GPU box).
usage: python tools/cli_trace.py [reads] [k]   -> KTB_FILE_TRACE lines of three runs of kmertools comp oligo"""
import os, subprocess, sys, tempfile, time
import numpy as np
m = int(sys.argv[1]) if len(sys.argv) > 1 else 217013
k = sys.argv[2] if len(sys.argv) > 2 else "5"
L = 150
rng = np.random.default_rng(1)
seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(m, L))]
rec = np.empty((m, 2 * L + 7), dtype=np.uint8)
rec[:, 0:3] = np.frombuffer(b"@r\n", dtype=np.uint8); rec[:, 3:3 + L] = seq
rec[:, 3 + L:6 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8); rec[:, 6 + L:6 + 2 * L] = ord("I"); rec[:, 6 + 2 * L] = ord("\n")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(root, "kmertools_b200", "bin", "kmertools")
with tempfile.TemporaryDirectory() as td:
    fq = os.path.join(td, "s.fq"); rec.tofile(fq)
    for i in range(3):
        t = time.perf_counter()
        r = subprocess.run([exe, "comp", "oligo", "-i", fq, "-o", os.path.join(td, "o.kmers"), "-k", k],
                           env=dict(os.environ, KTB_FILE_TRACE="1"), capture_output=True, text=True)
        print(f"run {i}: process wall {1e3 * (time.perf_counter() - t):.0f} ms |", r.stderr.strip())
