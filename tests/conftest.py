import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("ARTB_ASSETS", str(ROOT / "assets"))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def artb():
    import another_raytracer_b200 as A
    A.load()
    return A


@pytest.fixture(scope="session")
def gpu_ctx(artb):
    """One context for the whole GPU session. Fails loudly (no skip, no fallback) when CUDA is missing."""
    ctx = artb.Context(0, 1 << 20)
    yield ctx
    ctx.close()
