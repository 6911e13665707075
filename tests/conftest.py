import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden_traces(include_large=False):
    names = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".rvct.xz"))
    if not include_large:
        names = [n for n in names if not n.startswith("c2_4k")]
    return names


@pytest.fixture(scope="session")
def built():
    """Native pieces built (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True
