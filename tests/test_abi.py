"""The C-ABI library loads and exports every symbol include/locosim.h declares (no compute calls: CPU box)."""
import ctypes
import os
import re

import pytest

from helpers import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "locosim.h")).read()
    return sorted(set(re.findall(r"\b(locosim_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("locosim_create", "locosim_step", "locosim_reset", "locosim_destroy", "locosim_last_error"):
        assert s in syms


def test_shared_library_exports_every_declared_symbol():
    so = os.path.join(ROOT, "loco_mujoco_b200", "liblocosim_cuda.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(so)
    for s in declared_symbols():
        assert hasattr(lib, s), "missing export: " + s
    from loco_mujoco_b200 import engine
    assert sorted(engine.EXPORTED_SYMBOLS) == declared_symbols()


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from loco_mujoco_b200.engine import CudaEngine, EngineUnavailable
    with pytest.raises(EngineUnavailable):
        CudaEngine((None, None), (None, None), 4)
