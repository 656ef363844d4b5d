import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE = os.path.join(ROOT, "oracle")
if ORACLE not in sys.path:
    sys.path.insert(0, ORACLE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def tk():
    import tnrkit.jl_b200 as tk

    return tk


@pytest.fixture(scope="session")
def ctx(tk):
    return tk.default_context()
