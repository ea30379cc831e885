"""Builds libkmertools_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  cudart is linked
statically so the library depends on nothing but libcuda (the driver) at run time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIBDIR = HERE / "lib"
LIB = LIBDIR / "libkmertools_b200.so"
BIN = HERE / "bin" / "kmertools"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-ccbin", "/usr/bin/g++",
    "-cudart", "static",
    "--shared",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [d for d in CSRC.rglob("*") if d.is_file()] + [HERE.parent / "include" / "kmertools_b200.h", Path(__file__)]
    if not BIN.exists():
        return True
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    LIBDIR.mkdir(exist_ok=True)
    cmd = [nvcc(), *NVCC_FLAGS, *os.environ.get("KTB_NVCC_EXTRA", "").split()]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", str(LIB), *map(str, sources()), "-lz"]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    build_cli(env)
    return LIB


BIN = HERE / "bin" / "kmertools"


def build_cli(env=None) -> Path:
    """The `kmertools comp oligo` drop-in binary, linked against the shared library next to it."""
    BIN.parent.mkdir(exist_ok=True)
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-o", str(BIN), str(CSRC / "cli" / "main.cpp"),
           f"-L{LIBDIR}", "-lkmertools_b200", "-Wl,-rpath,$ORIGIN/../lib"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return BIN


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
