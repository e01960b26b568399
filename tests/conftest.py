import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: multi-minute CPU test, enabled with SPB_SLOW=1")


EMU_LIB = os.path.join(ROOT, "tests", "emu", "libspb_emu.so")


@pytest.fixture(scope="session")
def emu_lib():
    """Host emulation of the CUDA kernels (tests/emu): checks kernel LOGIC on the CPU-only tier."""
    src = [os.path.join(ROOT, "superpang_b200", "csrc", f) for f in ("spb.cu", "spb_kernels.cuh", "spb_platform.h")]
    if (not os.path.exists(EMU_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(EMU_LIB) for s in src):
        subprocess.check_call(["sh", os.path.join(ROOT, "tests", "emu", "build_emu.sh")])
    from superpang_b200 import _capi
    return _capi._declare(ctypes.CDLL(EMU_LIB))


@pytest.fixture(scope="session")
def cuda_engine_factory():
    """The product path: libsuperpang_b200.so on cuda:0.  Fails (not skips) if the library is missing."""
    from superpang_b200 import _capi

    def make():
        return _capi.Engine(device=0)
    return make
