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
    assert lib.pinn_version() == 102
    import ctypes as C
    out = C.c_void_p()
    rc = lib.pinn_plan_create(None, None, 0, 0, C.byref(out))
    assert rc == -1 and b"null" in lib.pinn_last_error()
    mlp = _capi.MlpDesc(2, 33, 3, 3)      # width not supported by any engine
    rc = lib.pinn_plan_create(C.byref(mlp), None, 0, 0, C.byref(out))
    assert rc == -1 and b"no engine" in lib.pinn_last_error()


def test_peer_memory_allreduce_rejects_bad_arguments_before_touching_cuda():
    _ensure_built()
    import ctypes as C
    lib = _capi.load()
    handle, ctx = (C.c_ubyte * 64)(), C.c_void_p()
    for world, rank, count in ((1, 0, 16), (9, 0, 16), (2, 2, 16), (2, 0, 0)):      # one rank, more than 8, rank out of range, empty
        assert lib.pinn_p2p_create(world, rank, 0, count, handle, C.byref(ctx)) == -1
        assert b"pinn_p2p_create" in lib.pinn_last_error()
    assert lib.pinn_p2p_allreduce_sum(None, None, 0, None) == -1
    assert lib.pinn_p2p_destroy(None) == 0


def test_struct_layout_matches_header():
    import ctypes as C
    # pinn_term_desc: 24 floats + conv + conv_k + rhs_scale (+4 pad) + ptr + 2 doubles + int64 + int32 (+4 pad)
    assert C.sizeof(_capi.TermDesc) == 96 + 12 + 4 + 8 + 16 + 8 + 8
    assert C.sizeof(_capi.PointSetDesc) == 8 + 8 + 8 + 8 * C.sizeof(_capi.TermDesc)
    assert C.sizeof(_capi.MlpDesc) == 16


def test_ctypes_structs_match_the_compiled_header(tmp_path):
    """sizeof / offsetof of every ABI struct as gcc lays out include/pinnstep.h == the ctypes mirror in _capi.py."""
    import ctypes as C
    src = tmp_path / "layout.c"
    src.write_text(r"""
#include <stdio.h>
#include <stddef.h>
#include "pinnstep.h"
int main(void) {
  printf("term %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(pinn_term_desc), offsetof(pinn_term_desc, conv),
         offsetof(pinn_term_desc, conv_k), offsetof(pinn_term_desc, rhs_scale), offsetof(pinn_term_desc, rhs_dev),
         offsetof(pinn_term_desc, weight), offsetof(pinn_term_desc, normalization), offsetof(pinn_term_desc, n_global),
         offsetof(pinn_term_desc, train), offsetof(pinn_term_desc, kind));
  printf("set %zu %zu %zu %zu %zu\n", sizeof(pinn_pointset_desc), offsetof(pinn_pointset_desc, n_local),
         offsetof(pinn_pointset_desc, n_terms), offsetof(pinn_pointset_desc, deriv_order), offsetof(pinn_pointset_desc, terms));
  printf("mlp %zu\n", sizeof(pinn_mlp_desc));
  printf("consts %d %d %d %d\n", PINN_MAX_OUT, PINN_MAX_CH, PINN_MAX_TERMS_PER_SET, PINN_VERSION);
  return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = dict((l.split()[0], [int(v) for v in l.split()[1:]]) for l in
               subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines())
    T, S = _capi.TermDesc, _capi.PointSetDesc
    assert out["term"] == [C.sizeof(T)] + [getattr(T, f).offset for f in
                                           ("conv", "conv_k", "rhs_scale", "rhs_dev", "weight", "normalization", "n_global", "train", "kind")]
    assert out["set"] == [C.sizeof(S), S.n_local.offset, S.n_terms.offset, S.deriv_order.offset, S.terms.offset]
    assert out["mlp"] == [C.sizeof(_capi.MlpDesc)]
    lib = (_ensure_built(), _capi.load())[1]
    assert out["consts"] == [_capi.MAX_OUT, _capi.MAX_CH, _capi.MAX_TERMS_PER_SET, lib.pinn_version()]
