import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_sessionfinish(session, exitstatus):
    """Achieved parity errors of this session (tests/helpers.py) -> gpurun_out/parity_r02.json (GPU sessions only)."""
    try:
        import helpers
        if any("gpu" in (r.get("test") or "") for r in helpers.PARITY_LOG):
            helpers.dump_parity_log(os.path.join(ROOT, "gpurun_out", "parity_r02.json"))
    except Exception:
        pass


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything the tests link against is (re)built once per session."""
    import __graft_entry__
    __graft_entry__.build()
