"""CPU: the C-ABI library builds/loads and exports every symbol include/rtdf.h declares (no compute calls)."""
import os
import re

from tests.util import ROOT, native


def _declared():
    src = open(os.path.join(ROOT, "include", "rtdf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rtdf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    nat = native()
    if not os.path.exists(nat.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("b", os.path.join(os.path.dirname(nat.LIB_PATH), "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = nat.load()
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in rtdf.h but not exported"
        assert name in nat.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(nat.SIGNATURES) == declared


def test_error_reporting_without_gpu():
    import ctypes
    nat = native()
    lib = nat.load()
    assert lib.rtdf_num_frames(64000) == 199
    assert lib.rtdf_num_frames(64600) == 201
    assert lib.rtdf_num_frames(16000) == 49
    assert lib.rtdf_num_frames(100) == 0
    ctx = ctypes.c_void_p()
    desc = nat.ModelDesc(backend=0, n_layers=25, precision=0)
    rc = lib.rtdf_create(ctypes.byref(ctx), 0, ctypes.byref(desc))
    assert rc != 0 and "at least 1 and at most 24" in nat.last_error()
