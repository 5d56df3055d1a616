"""The C-ABI library loads and exports exactly what include/pinnstep.h declares (no compute calls)."""
import os
import re
import subprocess

from pinns_fluid_dynamics_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()


def test_header_symbols_are_exported():
    _ensure_built()
    header = open(os.path.join(ROOT, "include", "pinnstep.h")).read()
    declared = set(re.findall(r"\b(pinn_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    nm = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (pinn_[a-z0-9_]+)", nm))
    assert declared <= exported, declared - exported


def test_library_loads_and_reports_errors_without_a_gpu():
    _ensure_built()
    lib = _capi.load()
    assert lib.pinn_version() == 101
    import ctypes as C
    out = C.c_void_p()
    rc = lib.pinn_plan_create(None, None, 0, 0, C.byref(out))
    assert rc == -1 and b"null" in lib.pinn_last_error()
    mlp = _capi.MlpDesc(2, 33, 3, 3)      # width not supported by any engine
    rc = lib.pinn_plan_create(C.byref(mlp), None, 0, 0, C.byref(out))
    assert rc == -1 and b"no engine" in lib.pinn_last_error()


def test_struct_layout_matches_header():
    import ctypes as C
    # pinn_term_desc: 24 floats + conv + conv_k + rhs_scale (+4 pad) + ptr + 2 doubles + int64 + int32 (+4 pad)
    assert C.sizeof(_capi.TermDesc) == 96 + 12 + 4 + 8 + 16 + 8 + 8
    assert C.sizeof(_capi.PointSetDesc) == 8 + 8 + 8 + 8 * C.sizeof(_capi.TermDesc)
    assert C.sizeof(_capi.MlpDesc) == 16
