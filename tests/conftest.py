"""pytest configuration: `gpu` marker for tests that need a B200, everything else runs on CPU."""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """make sure the oracle / generator / hostcheck libraries exist (built by __graft_entry__.build())"""
    need = [ROOT / "oracle" / "liblvi_oracle.so", ROOT / "tools" / "synth" / "liblvi_synth.so", ROOT / "lvi_exc_b200" / "lib" / "liblvi_exc_b200.so"]
    if not all(p.exists() for p in need):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session")
def cuda_backend():
    from lvi_exc_b200.backend import CudaBackend
    b = CudaBackend(0)
    yield b
    b.close()
