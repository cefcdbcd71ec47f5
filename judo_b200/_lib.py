"""ctypes binding of libb200mpc.so (include/b200mpc.h).  Fails loudly when the CUDA library is missing:
there is NO CPU fallback on the product path."""

from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200mpc.so")

_dp = ctypes.POINTER(ctypes.c_double)
_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p
_i = ctypes.c_int
_d = ctypes.c_double


class Dims(ctypes.Structure):
    _fields_ = [("nq", _i), ("nv", _i), ("nu", _i), ("nsensordata", _i), ("n_cost_params", _i)]


# symbol -> (restype, argtypes); mirrors include/b200mpc.h one to one.  Array arguments are declared void*: callers pass raw
# addresses (ndarray.ctypes.data), which costs ~1 us per argument instead of ~2.5 us for data_as(POINTER(c_double)) — that
# difference is visible in a 100 us plan step.
SIGNATURES = {
    "b200mpc_create": (_i, [ctypes.POINTER(_vp), _i, _vp, ctypes.c_size_t, _i, _i]),
    "b200mpc_destroy": (None, [_vp]),
    "b200mpc_last_error": (ctypes.c_char_p, [_vp]),
    "b200mpc_get_dims": (_i, [_vp, ctypes.POINTER(Dims)]),
    "b200mpc_update": (_i, [_vp, _i]),
    "b200mpc_num_rollouts": (_i, [_vp]),
    "b200mpc_rollout": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp]),
    "b200mpc_plan_costs": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "b200mpc_reward": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "b200mpc_reward_sensors": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "b200mpc_update_mppi": (_i, [_vp, _vp, _vp, _i, _i, _d, _vp]),
    "b200mpc_update_cem": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _d, _vp, _vp]),
    "b200mpc_update_ps": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "b200mpc_plan_step": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "b200mpc_plan_step_sampled": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, _i, _vp, ctypes.c_ulonglong, ctypes.c_ulonglong, _i,
                                       _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "b200mpc_plan_costs_dev": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "b200mpc_plan_step_dev": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200mpc_rollout_dev": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "b200mpc_mppi_partial_dev": (_i, [_vp, _vp, _vp, _i, _i, _d, _vp, _vp]),
    "b200mpc_mppi_combine_dev": (_i, [_vp, _vp, _i, _i, _d, _vp, _vp]),
    "b200mpc_topk_partial_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "b200mpc_topk_combine_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _d, _vp, _vp, _vp, _vp]),
    "b200mpc_exchange_create": (_i, [_vp, _i, _i, _vp]),
    "b200mpc_exchange_open": (_i, [_vp, _vp]),
    "b200mpc_exchange_align_dev": (_i, [_vp, _vp]),
    "b200mpc_exchange_buffer": (_i, [_vp, _vp]),
    "b200mpc_exchange_open_local": (_i, [_vp, _vp]),
    "b200mpc_exchange_stamps": (_i, [_vp, _vp]),
    "b200mpc_exchange_align_stamp": (_i, [_vp, _vp]),
    "b200mpc_launch_count": (ctypes.c_longlong, [_vp]),
    "b200mpc_fp64_peak": (_i, [_i, ctypes.POINTER(_d)]),
    "b200mpc_contact_overflows": (ctypes.c_longlong, [_vp]),
    "b200mpc_set_trace_capture": (_i, [_vp, _i]),
    "b200mpc_trace_width": (_i, [_vp]),
    "b200mpc_elite_traces": (_i, [_vp, _vp, _i, _i, _vp]),
    "b200mpc_controller_step": (_i, [_vp, _vp]),
    "b200mpc_last_candidates": (_i, [_vp, _vp, _i, _i]),
    "b200mpc_controller_speculation": (_i, [_vp, _vp, _vp, ctypes.c_size_t]),
    "b200mpc_group_create": (_i, [ctypes.POINTER(_vp), _i, _vp, ctypes.c_size_t, _vp, _i, _i]),
    "b200mpc_group_destroy": (None, [_vp]),
    "b200mpc_group_last_error": (ctypes.c_char_p, [_vp]),
    "b200mpc_group_size": (_i, [_vp]),
    "b200mpc_group_handle": (_vp, [_vp, _i]),
    "b200mpc_group_update": (_i, [_vp, _i]),
    "b200mpc_group_num_rollouts": (_i, [_vp]),
    "b200mpc_group_rollout": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp]),
    "b200mpc_group_plan_step": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "b200mpc_group_controller_step": (_i, [_vp, _vp]),
    "b200mpc_group_controller_speculation": (_i, [_vp, _vp, _vp, ctypes.c_size_t]),
    "b200mpc_group_last_candidates": (_i, [_vp, _vp, _i, _i]),
    "b200mpc_group_set_trace_capture": (_i, [_vp, _i]),
    "b200mpc_group_elite_traces": (_i, [_vp, _vp, _i, _i, _vp]),
    "b200mpc_group_launch_count": (ctypes.c_longlong, [_vp]),
    "b200mpc_group_contact_overflows": (ctypes.c_longlong, [_vp]),
    "b200mpc_legacy_normals": (_i, [_vp, _vp, _vp, ctypes.c_size_t]),
    "b200mpc_spline_basis": (_i, [_i, _vp, _i, _vp, _i, _vp]),
}


class StepRequest(ctypes.Structure):
    """b200mpc_step_request (include/b200mpc.h)."""

    _fields_ = [("N", _i), ("K", _i), ("H", _i), ("optimizer", _i), ("spline_order", _i), ("n_elite", _i), ("phase", _i), ("n_head", _i),
                ("has_tail", _i), ("n_trace_sensors", _i), ("speculate", _i), ("use_speculated", _i), ("time", _d), ("dt", _d), ("head", _d * 2), ("tail", _d),
                ("knot_times", _vp), ("x0", _vp), ("nominal", _vp), ("sigma", _vp), ("lo", _vp), ("hi", _vp), ("cost_params", _vp),
                ("opt_params", _vp), ("trace_cols", _vp), ("mt_key", _vp), ("mt_pos", _vp),
                ("nominal_out", _vp), ("sigma_out", _vp), ("rewards", _vp), ("elite_idx", _vp), ("traces", _vp), ("basis_out", _vp),
                ("knots_out", _vp)]

_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m judo_b200.build` (nvcc, sm_100a). "
                "judo_b200 has no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library ever diverge
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
