import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def synth():
    import helpers
    return helpers.load_synth()


@pytest.fixture(scope="session")
def oracle():
    import helpers
    return helpers.load_oracle()


@pytest.fixture(scope="session")
def emu():
    """CPU debugging build of the CUDA sources (tests/emu) — test infrastructure, not the product."""
    import helpers
    return helpers.load_emu()


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library.  GPU tests fail (not skip) when it is missing or no device is present."""
    import helpers
    return helpers.D.load()
