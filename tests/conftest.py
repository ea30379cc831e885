import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The shared library and the CLI are build artefacts (git-ignored): make sure they exist and are current
    # before any test imports them.  nvcc cross-compiles sm_100a without a GPU; this is a no-op when up to date.
    from kmertools_b200 import build as kb
    try:
        kb.build()
    except Exception as exc:  # e.g. a box without nvcc: use what travelled with the snapshot
        if not kb.LIB.exists():
            raise
        print(f"[conftest] using the prebuilt library ({exc})")


@pytest.fixture(scope="session")
def golden():
    return GOLDEN
