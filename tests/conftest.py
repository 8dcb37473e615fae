import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ob():
    """The CPU oracle (test infrastructure): built on first use."""
    from oracle import binding

    binding.lib()
    return binding


@pytest.fixture(scope="session")
def synth():
    from msckf_stereo_c_b200 import synth as s

    s.lib()
    return s


def copy_cfg(cfg, **kw):
    from msckf_stereo_c_b200 import abi

    c = abi.Config.from_buffer_copy(bytes(cfg))
    for k, v in kw.items():
        setattr(c, k, v)
    return c
