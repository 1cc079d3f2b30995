import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    import helpers
    helpers.build_oracle()
    return helpers.oracle()


@pytest.fixture(scope="session")
def emu_lib():
    import helpers
    return helpers.emu()


@pytest.fixture(params=["per-block", "per-input"])
def staged_form(request, emu_lib):
    """Both forms of the staged zstd decoder under the emulator: 'per-block' (one lane / quad per (input, block), up to 32
    blocks per input: what the device picks for few inputs -- and the emulator's default, since test batches are small) and
    'per-input' (at most four blocks per input: what it picks for many)."""
    emu_lib.emu_set_dec_maxb(4 if request.param == "per-input" else 0)
    yield request.param
    emu_lib.emu_set_dec_maxb(0)
