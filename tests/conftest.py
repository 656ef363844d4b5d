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


def pytest_addoption(parser):
    parser.addoption("--emulate-abi", action="store_true", default=False,
                     help="DEVELOPMENT AID: run `-m gpu` tests of the block-sparse HOST layer on "
                          "tests/abi_emulator.py (numpy) instead of libtnrcuda, so that a new "
                          "device test can be checked on the CPU before it is committed.  Tests "
                          "that call dense step entries or raw kernels cannot run this way.")


@pytest.fixture(autouse=True)
def _maybe_emulated_abi(request, monkeypatch):
    if request.config.getoption("--emulate-abi"):
        from abi_emulator import EmulatedContext
        from tnrkit.jl_b200 import _lib

        monkeypatch.setattr(_lib, "_default_ctx", EmulatedContext())
    yield


@pytest.fixture(scope="session")
def tk():
    import tnrkit.jl_b200 as tk

    return tk


@pytest.fixture(scope="session")
def ctx(tk):
    return tk.default_context()
