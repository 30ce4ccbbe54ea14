import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libludwig_oracle.so")
CUDA_LIB = os.path.join(ROOT, "open_ludwig_b200", "csrc", "libludwig_b200.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU parity oracle (test infrastructure).  Built on demand with oracle/Makefile."""
    if not os.path.exists(ORACLE_LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return ORACLE_LIB


@pytest.fixture(scope="session")
def cuda_lib():
    assert os.path.exists(CUDA_LIB), "libludwig_b200.so missing: run __graft_entry__.build() first"
    return CUDA_LIB
