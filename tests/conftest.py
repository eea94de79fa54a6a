import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emu():
    """ctypes handle on the host emulation of the per-cell stage functions (test infrastructure)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "host_emu")
    so = os.path.join(d, "libsol_emu.so")
    src = os.path.join(d, "emu.cpp")
    hdrs = [os.path.join(ROOT, "solver_in_the_loop_b200", "csrc", h) for h in ("sol_cells.cuh", "sol_direct_host.h")]
    if (not os.path.exists(so)) or os.path.getmtime(so) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-o", so, src])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
