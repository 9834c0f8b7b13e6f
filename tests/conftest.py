import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The CPU suite needs the oracle and the synthetic generator; both are plain gcc builds."""
    from oracle import pyoracle
    pyoracle.lib()
    import raymarchdenoisercuda_b200.build as b
    if not (os.path.exists(os.path.join(ROOT, "raymarchdenoisercuda_b200", "librmd_b200.so"))
            and os.path.exists(os.path.join(ROOT, "raymarchdenoisercuda_b200", "librmd_synth.so"))):
        b.build()
    yield
