import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # pytest-timeout registers this itself when installed; declare it so the suite also collects cleanly without the plugin
    config.addinivalue_line("markers", "timeout(seconds): per-test time limit (pytest-timeout)")


@pytest.fixture(scope="session")
def engine():
    """One librsvdb context on cuda:0.  Fails (does not skip) when the CUDA library cannot be loaded or no GPU exists:
    a -m gpu run that cannot reach the product is an error, not a pass."""
    from rsvd_kamaneh_raganato_terrana_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="session")
def oracle():
    from oracle import rsvd_oracle
    rsvd_oracle.build()
    return rsvd_oracle
