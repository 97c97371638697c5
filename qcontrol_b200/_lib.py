"""ctypes binding of the C ABI declared in include/qcontrol_b200.h.

The CUDA library is the product: if it is missing or does not load, importing
this module raises -- there is no CPU fallback anywhere in qcontrol_b200.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QCONTROL_B200_LIB") or os.path.join(_HERE, "lib", "libqcontrol_b200.so")

ABI_VERSION = 1

# every symbol include/qcontrol_b200.h declares
SYMBOLS = (
    "qc_abi_version", "qc_last_error", "qc_create", "qc_destroy", "qc_synchronize", "qc_timer_start",
    "qc_timer_stop", "qc_launch_count", "qc_measure_fp64_peak", "qc_mem_info", "qc_sm_count", "qc_pool_reserve", "qc_plan_create", "qc_plan_destroy",
    "qc_plan_footprint",
    "qc_set_static", "qc_set_poly", "qc_set_poly_orders", "qc_set_state", "qc_get_state", "qc_set_step_scalars", "qc_set_control",
    "qc_sweep", "qc_prop_step", "qc_apply_hamil", "qc_set_instr_grid", "qc_instrument", "qc_set_local_control", "qc_get_local_state", "qc_set_local_state", "qc_get_fields", "qc_field_integral", "qc_get_norms", "qc_snapshot_set", "qc_snapshot_load", "qc_goal_overlap", "qc_krotov_terminal",
    "qc_get_traj", "qc_record_state", "qc_propagate_host",
    "qc_plan2d_create", "qc_plan2d_destroy", "qc_plan2d_set_static", "qc_plan2d_set_poly", "qc_plan2d_set_state",
    "qc_plan2d_get_state", "qc_plan2d_sweep", "qc_plan2d_get_norms",
)


class PlanDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("hamil", "np", "nlevs", "nb", "nch", "batch", "nt_max", "hf_hide", "store_traj", "reserved")]


class SweepArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("poly_slot", "field_mode", "l_first", "nsteps", "renorm", "traj_read", "traj_write",
                 "reversed_E", "write_E_base")] + [("reserved", C.c_int32 * 3)]


class QcError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("qcontrol_b200 error %d: %s" % (code, message))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "qcontrol_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C qcontrol_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    i32 = C.c_int32
    sig = {
        "qc_abi_version": (i32, []),
        "qc_last_error": (C.c_char_p, []),
        "qc_create": (i32, [i32, vp, C.POINTER(vp)]),
        "qc_destroy": (i32, [vp]),
        "qc_synchronize": (i32, [vp]),
        "qc_timer_start": (i32, [vp]),
        "qc_timer_stop": (i32, [vp, dp]),
        "qc_launch_count": (C.c_int64, [vp]),
        "qc_measure_fp64_peak": (i32, [vp, dp]),
        "qc_mem_info": (i32, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        "qc_sm_count": (i32, [vp, C.POINTER(C.c_int32)]),
        "qc_pool_reserve": (i32, [vp, C.c_int64]),
        "qc_plan_footprint": (i32, [C.POINTER(PlanDesc), C.POINTER(C.c_int64)]),
        "qc_plan_create": (i32, [vp, C.POINTER(PlanDesc), C.POINTER(vp)]),
        "qc_plan_destroy": (i32, [vp]),
        "qc_set_static": (i32, [vp, dp, i32, dp, dp, i32, dp, i32]),
        "qc_set_poly": (i32, [vp, i32, dp, dp, dp, i32]),
        "qc_set_poly_orders": (i32, [vp, i32, C.POINTER(C.c_int32), i32]),
        "qc_set_state": (i32, [vp, dp]),
        "qc_get_state": (i32, [vp, dp]),
        "qc_set_step_scalars": (i32, [vp, i32, dp, i32]),
        "qc_set_control": (i32, [vp, dp, i32, dp]),
        "qc_sweep": (i32, [vp, C.POINTER(SweepArgs)]),
        "qc_prop_step": (i32, [vp, i32, dp]),
        "qc_apply_hamil": (i32, [vp, i32, i32, dp, C.c_double, dp]),
        "qc_set_instr_grid": (i32, [vp, dp, dp]),
        "qc_instrument": (i32, [vp, dp, C.c_double, i32, i32, dp]),
        "qc_set_local_control": (i32, [vp, dp, dp, dp]),
        "qc_get_local_state": (i32, [vp, dp]),
        "qc_set_local_state": (i32, [vp, dp]),
        "qc_plan2d_create": (i32, [vp, i32, i32, i32, C.POINTER(vp)]),
        "qc_plan2d_destroy": (i32, [vp]),
        "qc_plan2d_set_static": (i32, [vp, dp, i32, dp, dp, C.c_double]),
        "qc_plan2d_set_poly": (i32, [vp, dp, dp, dp]),
        "qc_plan2d_set_state": (i32, [vp, dp]),
        "qc_plan2d_get_state": (i32, [vp, dp]),
        "qc_plan2d_sweep": (i32, [vp, i32, i32]),
        "qc_plan2d_get_norms": (i32, [vp, dp]),
        "qc_get_fields": (i32, [vp, dp]),
        "qc_field_integral": (i32, [vp, dp]),
        "qc_get_norms": (i32, [vp, dp]),
        "qc_snapshot_set": (i32, [vp, i32, dp]),
        "qc_snapshot_load": (i32, [vp, i32]),
        "qc_goal_overlap": (i32, [vp, i32, dp]),
        "qc_krotov_terminal": (i32, [vp, i32, dp, dp, i32, i32]),
        "qc_get_traj": (i32, [vp, i32, i32, dp]),
        "qc_record_state": (i32, [vp, i32, i32, i32]),
        "qc_propagate_host": (i32, [vp, C.POINTER(SweepArgs), dp, dp]),
    }
    assert set(sig) == set(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    got = lib.qc_abi_version()
    if got != ABI_VERSION:
        raise ImportError("qcontrol_b200: library ABI %d, binding expects %d" % (got, ABI_VERSION))
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise QcError(rc, lib.qc_last_error().decode("utf-8", "replace"))
