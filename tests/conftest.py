import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The library under test must be the build of the sources in the tree (content hash in ce_version()): rebuild a
    stale one before any test maps it (dlopen would keep serving the old mapping to the whole session)."""
    from codec_eval_b200 import build

    try:
        build.ensure_built()
    except Exception as e:   # no nvcc: the tests that load the library then fail loudly on their own
        sys.stderr.write(f"conftest: could not (re)build libce_gpu.so: {e}\n")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gpu():
    """One GpuMetrics context for the whole session; fails loudly without the CUDA library / a GPU."""
    from codec_eval_b200.metrics import GpuMetrics

    ctx = GpuMetrics(0, workspace_bytes=6 << 30)
    yield ctx
    ctx.close()
