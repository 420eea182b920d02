"""The C-ABI library loads and exports every symbol include/stretchsim.h declares (CPU only,
no compute calls)."""
import ctypes
import os
import re

from stretch_mujoco_b200 import engine

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def _declared():
    src = open(os.path.join(ROOT, "include", "stretchsim.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(engine.LIB_PATH), "libstretchsim.so not built (run __graft_entry__.build())"
    L = ctypes.CDLL(engine.LIB_PATH)
    decl = _declared()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(L, name), f"{name} declared in stretchsim.h but not exported"
    assert sorted(engine.EXPORTS) == decl


def test_version_and_error_strings():
    L = engine.lib()
    assert b"sm_100a" in L.ss_version()
    assert L.ss_name2id(None, 0, b"x") == -1


def test_no_cpu_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.StretchSimError):
        engine.DeviceModel(b"SSMBLOB1" + b"\0" * 64, 0)
