"""pytest configuration: `gpu` marker, import paths, and in-tree builds of the two native
libraries (libotmb.so = the product, CUDA; libotmb_oracle.so = the CPU checker)."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu")
    # build what is missing (nvcc cross-compiles without a GPU; the .so files travel to the GPU box)
    import otmb_b200.build as b
    b.build()
    from oracle import oracle
    oracle.build()


@pytest.fixture(scope="session")
def ctx():
    import otmb_b200
    return otmb_b200.default_context(0)
