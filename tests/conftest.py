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


@pytest.fixture(scope="session")
def emu():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "liblocosim_emu.so")
    csrc = os.path.join(ROOT, "loco_mujoco_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in ("locosim_emu.cpp", "locosim_core.cuh", "locosim_host.h", "locosim_config.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-o", so, srcs[0], "-lm"])
    import ctypes
    lib = ctypes.CDLL(so)
    lib.emu_create.restype = ctypes.c_void_p
    lib.emu_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    lib.emu_destroy.argtypes = [ctypes.c_void_p]
    for f in ("emu_reset", "emu_step", "emu_get_state"):
        getattr(lib, f).restype = None
    lib.emu_reset.argtypes = lib.emu_get_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.emu_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return lib
