import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """ctypes handle of the CPU oracle (test infrastructure). Built on demand with oracle/Makefile."""
    so = os.path.join(ROOT, "oracle", "liblocosim_ref.so")
    src = os.path.join(ROOT, "oracle", "locosim_ref.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    import oracle_binding
    return oracle_binding.load(so)


@pytest.fixture(scope="session")
def bundled_only():
    """Force the package to use its bundled assets (what the GPU box sees: no /root/reference there)."""
    os.environ["LOCO_MUJOCO_B200_FORCE_BUNDLED"] = "1"
    yield
    os.environ.pop("LOCO_MUJOCO_B200_FORCE_BUNDLED", None)
