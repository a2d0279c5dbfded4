"""ctypes binding of libdqmc_b200.so -- exactly the symbols of include/dqmc_b200.h.

There is no CPU fallback: if the shared library is missing it is built with nvcc; if that
fails, or no B200 is present when a context is created, the call raises.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_HEADER = _HERE.parent / "include" / "dqmc_b200.h"
_lib = None

dp = C.POINTER(C.c_double)
i8p = C.POINTER(C.c_int8)
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)


class Desc(C.Structure):
    _fields_ = [("n_sites", C.c_int32), ("n_slices", C.c_int32), ("field_kind", C.c_int32),
                ("n_chains", C.c_int32), ("n_ranges", C.c_int32), ("range_first", i32p),
                ("range_last", i32p), ("alpha", C.c_double), ("hopping_exp_squared", dp),
                ("hopping_exp_inv_squared", dp), ("hopping_exp", dp), ("hopping_exp_inv", dp),
                ("check_sign_problem", C.c_int32), ("check_propagation_error", C.c_int32),
                ("seed", C.c_uint64), ("chain_offset", C.c_int64), ("device", C.c_int32),
                ("delay_block", C.c_int32), ("update_variant", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("neg_count", C.c_int64), ("neg_sumlog10", C.c_double), ("neg_min", C.c_double),
                ("neg_max", C.c_double), ("prop_count", C.c_int64), ("prop_sumlog10", C.c_double),
                ("prop_min", C.c_double), ("prop_max", C.c_double)]


def declared_symbols() -> list[str]:
    """Every function include/dqmc_b200.h declares."""
    text = _HEADER.read_text()
    return sorted(set(re.findall(r"\b(dqmc_[a-z0-9_]+)\s*\(", text)))


def library_path() -> Path:
    return _HERE / "libdqmc_b200.so"


def load():
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    so = _build.build()          # no-op when up to date; raises if nvcc fails
    L = C.CDLL(str(so))
    vp = C.c_void_p
    i32, i64 = C.c_int32, C.c_int64
    sig = {
        "dqmc_create": (i32, [C.POINTER(Desc), C.POINTER(vp)]),
        "dqmc_destroy": (i32, [vp]),
        "dqmc_last_error": (C.c_char_p, [vp]),
        "dqmc_device_count": (i32, []),
        "dqmc_set_conf": (i32, [vp, i32, i32, i8p]),
        "dqmc_get_conf": (i32, [vp, i32, i32, i8p]),
        "dqmc_get_conf_packed": (i32, [vp, i32, i32, C.POINTER(C.c_uint64)]),
        "dqmc_set_conf_packed": (i32, [vp, i32, i32, C.POINTER(C.c_uint64)]),
        "dqmc_build_stack": (i32, [vp]),
        "dqmc_forward_build_stack": (i32, [vp]),
        "dqmc_propagate": (i32, [vp, i32]),
        "dqmc_get_state": (i32, [vp, i32p]),
        "dqmc_sweep": (i32, [vp, i32, dp, i64p]),
        "dqmc_sweep_traced": (i32, [vp, dp, u8p, dp, u8p, i64p]),
        "dqmc_sweep_spatial": (i32, [vp, dp, u8p, dp, u8p, i64p]),
        "dqmc_set_sweep_index": (i32, [vp, i64]),
        "dqmc_get_greens": (i32, [vp, i32, i32, dp]),
        "dqmc_set_greens": (i32, [vp, i32, i32, dp]),
        "dqmc_get_measured_greens": (i32, [vp, i32, i32, dp]),
        "dqmc_calculate_greens_at": (i32, [vp, i32, i32, dp]),
        "dqmc_get_stats": (i32, [vp, i32, i32, C.POINTER(Stats)]),
        "dqmc_get_stack_array": (i32, [vp, i32, i32, i32, dp]),
        "dqmc_ut_build_stack": (i32, [vp]),
        "dqmc_ut_lazy_build": (i32, [vp, i32, i32]),
        "dqmc_ut_greens": (i32, [vp, i32, i32, i32, dp]),
        "dqmc_ut_get_stack_array": (i32, [vp, i32, i32, i32, dp]),
        "dqmc_cgi_begin": (i32, [vp, i32, i32, i32, i32]),
        "dqmc_cgi_next": (i32, [vp, i32p, dp, dp, dp]),
        "dqmc_global_update": (i32, [vp, i8p, dp, i32, i64p, dp]),
        "dqmc_set_global_update_index": (i32, [vp, i64]),
        "dqmc_comm_unique_id": (i32, [u8p]),
        "dqmc_comm_init": (i32, [vp, i32, i32, u8p]),
        "dqmc_comm_destroy": (i32, [vp]),
        "dqmc_set_lattice": (i32, [vp, i32, i32, i32p, dp, C.c_double]),
        "dqmc_measurement_layout": (i32, [vp, i32p]),
        "dqmc_measure_equal_time": (i32, [vp]),
        "dqmc_measure_time_integral": (i32, [vp, i32, i32, C.c_double]),
        "dqmc_get_measurements": (i32, [vp, i32, i32, dp]),
        "dqmc_measurement_buffer": (i32, [vp, C.POINTER(vp), i64p]),
        "dqmc_get_measurement_stats": (i32, [vp, dp, dp, dp]),
        "dqmc_binning_levels": (i32, []),
        "dqmc_get_measurement_binning": (i32, [vp, dp, dp, dp]),
        "dqmc_measurement_binning_buffer": (i32, [vp, C.POINTER(vp), i64p]),
        "dqmc_accumulate_greens": (i32, [vp]),
        "dqmc_observable_buffer": (i32, [vp, C.POINTER(vp), i64p]),
        "dqmc_reduce_observables": (i32, [vp, vp]),
        "dqmc_get_observables": (i32, [vp, dp, dp, dp]),
        "dqmc_op_vmul": (i32, [i32, i32, i32, i32, i32, dp, dp, dp]),
        "dqmc_op_udt": (i32, [i32, i32, i32, i32, dp, dp, dp, dp, i64p]),
        "dqmc_op_rdivp": (i32, [i32, i32, i32, dp, dp, i64p]),
        "dqmc_op_calculate_greens": (i32, [i32, i32, i32, dp, dp, dp, dp, dp, dp, dp]),
        "dqmc_op_multiply_slice_matrix": (i32, [vp, i32, i32, dp]),
        "dqmc_op_wrap_greens": (i32, [vp, i32, i32, dp]),
        "dqmc_get_stream": (i32, [vp, C.POINTER(vp)]),
        "dqmc_profile": (i32, [vp, i32]),
        "dqmc_profile_report": (i32, [vp, dp, i64p]),
        "dqmc_kernel_launches": (i64, [vp]),
        "dqmc_max_sites": (i32, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)      # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
