import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The CPU oracle is cheap to (re)build; the CUDA library must already be there (build())."""
    import oracle

    oracle.build()
    lib = os.path.join(ROOT, "cogdl_b200", "lib", "libcogdl_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__

        __graft_entry__.build()
