import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def hostemu():
    """tests/hostemu/libhostemu.so: the device headers compiled for the CPU (test harness only)."""
    import ctypes
    so = os.path.join(ROOT, "tests", "hostemu", "libhostemu.so")
    src = os.path.join(ROOT, "tests", "hostemu", "hostemu.cpp")
    hdrs = [os.path.join(ROOT, "mathlib_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "mathlib_b200", "csrc"))
            if f.endswith((".cuh", ".h"))]
    newest = max(os.path.getmtime(p) for p in [src] + hdrs)
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)
