"""ctypes binding of libpinnstep.so (include/pinnstep.h).  No CPU fallback: if the library cannot be
loaded every compute entry point raises ``PinnLibraryError``."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

MAX_OUT, MAX_CH, MAX_DIM, MAX_TERMS_PER_SET = 4, 6, 3, 8

LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
LIB_PATH = os.environ.get("PINN_LIBPINNSTEP", os.path.join(LIB_DIR, "libpinnstep.so"))   # override: development builds only

# every symbol include/pinnstep.h declares (tests/test_capi_symbols.py checks the header against this)
EXPORTED_SYMBOLS = (
    "pinn_version", "pinn_last_error", "pinn_plan_create", "pinn_plan_destroy", "pinn_plan_param_count",
    "pinn_plan_term_count", "pinn_plan_workspace_bytes", "pinn_plan_engine", "pinn_plan_last_launch_count",
    "pinn_plan_set_rhs", "pinn_plan_enable_timing", "pinn_plan_kernel_time_ms", "pinn_loss_and_grad", "pinn_loss", "pinn_forward", "pinn_nccl_unique_id",
    "pinn_comm_create", "pinn_comm_destroy", "pinn_allreduce_sum", "pinn_p2p_create", "pinn_p2p_connect", "pinn_p2p_allreduce_sum",
    "pinn_p2p_status", "pinn_p2p_destroy", "pinn_adam_step", "pinn_adam_step_dev",
    "pinn_bfgs_identity", "pinn_bfgs_trial", "pinn_bfgs_trial_dev", "pinn_bfgs_eval", "pinn_bfgs_direction", "pinn_bfgs_accept_update",
)


class PinnLibraryError(RuntimeError):
    pass


class MlpDesc(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("width", C.c_int32), ("n_hidden", C.c_int32), ("out_dim", C.c_int32)]


class TermDesc(C.Structure):
    _fields_ = [
        ("coef", (C.c_float * MAX_CH) * MAX_OUT),
        ("conv", C.c_float),
        ("conv_k", C.c_int32),
        ("rhs_scale", C.c_float),
        ("rhs_dev", C.c_void_p),
        ("weight", C.c_double),
        ("normalization", C.c_double),
        ("n_global", C.c_int64),
        ("train", C.c_int32),
        ("kind", C.c_int32),
    ]


class PointSetDesc(C.Structure):
    _fields_ = [
        ("points_dev", C.c_void_p),
        ("n_local", C.c_int64),
        ("n_terms", C.c_int32),
        ("deriv_order", C.c_int32),
        ("terms", TermDesc * MAX_TERMS_PER_SET),
    ]


_lib: Optional[C.CDLL] = None


def load(path: Optional[str] = None) -> C.CDLL:
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise PinnLibraryError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    try:
        lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    except OSError as e:  # pragma: no cover - depends on the box
        raise PinnLibraryError(f"cannot load {p}: {e}") from e
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.pinn_version.restype = C.c_int
    lib.pinn_last_error.restype = C.c_char_p
    lib.pinn_plan_create.argtypes = [C.POINTER(MlpDesc), C.POINTER(PointSetDesc), i32, i32, C.POINTER(vp)]
    lib.pinn_plan_destroy.argtypes = [vp]
    lib.pinn_plan_param_count.argtypes = [vp]; lib.pinn_plan_param_count.restype = i64
    lib.pinn_plan_term_count.argtypes = [vp]; lib.pinn_plan_term_count.restype = i32
    lib.pinn_plan_workspace_bytes.argtypes = [vp]; lib.pinn_plan_workspace_bytes.restype = C.c_size_t
    lib.pinn_plan_engine.argtypes = [vp]; lib.pinn_plan_engine.restype = C.c_char_p
    lib.pinn_plan_last_launch_count.argtypes = [vp]; lib.pinn_plan_last_launch_count.restype = i32
    lib.pinn_plan_set_rhs.argtypes = [vp, i32, i32, vp]
    lib.pinn_plan_enable_timing.argtypes = [vp, i32]
    lib.pinn_plan_kernel_time_ms.argtypes = [vp, i32, C.POINTER(C.c_float)]
    lib.pinn_loss_and_grad.argtypes = [vp, vp, vp, vp]
    lib.pinn_loss.argtypes = [vp, vp, vp, vp]
    lib.pinn_forward.argtypes = [C.POINTER(MlpDesc), vp, vp, i64, vp, i32, vp]
    lib.pinn_nccl_unique_id.argtypes = [vp]
    lib.pinn_comm_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
    lib.pinn_comm_destroy.argtypes = [vp]
    lib.pinn_allreduce_sum.argtypes = [vp, vp, i64, vp]
    lib.pinn_p2p_create.argtypes = [i32, i32, i32, i64, vp, C.POINTER(vp)]
    lib.pinn_p2p_connect.argtypes = [vp, vp]
    lib.pinn_p2p_allreduce_sum.argtypes = [vp, vp, i64, vp]
    lib.pinn_p2p_status.argtypes = [vp, C.POINTER(i32)]
    lib.pinn_p2p_destroy.argtypes = [vp]
    lib.pinn_adam_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, i64, vp]
    lib.pinn_adam_step_dev.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, vp, vp]
    lib.pinn_bfgs_identity.argtypes = [vp, i64, vp]
    lib.pinn_bfgs_trial.argtypes = [vp, vp, C.c_double, vp, vp, i64, vp]
    lib.pinn_bfgs_trial_dev.argtypes = [vp, vp, vp, vp, vp, i64, vp]
    lib.pinn_bfgs_eval.argtypes = [vp, vp, vp, i32, vp, vp, vp, i64, vp]
    lib.pinn_bfgs_direction.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.pinn_bfgs_accept_update.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i64, vp]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("pinn_version", "pinn_last_error", "pinn_plan_param_count", "pinn_plan_term_count",
                        "pinn_plan_workspace_bytes", "pinn_plan_engine", "pinn_plan_last_launch_count"):
            fn.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pinn_last_error()
        raise PinnLibraryError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
