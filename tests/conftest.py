import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


import pytest  # noqa: E402


@pytest.fixture
def emu(monkeypatch):
    """Route MultiModN through the host-compiled kernel emulator (tests/emu) — CPU tests only."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    from emu_backend import get_emu_lib
    from multimodn_b200 import MultiModN
    lib = get_emu_lib()
    monkeypatch.setattr(MultiModN, "_lib_factory", staticmethod(lambda: lib))
    return lib
