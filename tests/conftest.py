import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    # The product refuses to run without its CUDA library (no CPU fallback).  In a checkout where it has not been
    # built yet, build it once here (nvcc cross-compiles without a GPU) — same as __graft_entry__.build().
    from fqtk_b200 import build as _build

    if _build.is_stale():
        _build.build()


@pytest.fixture(scope="session")
def kats():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as fh:
        return json.load(fh)
