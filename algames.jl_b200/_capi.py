"""ctypes binding of the C ABI in include/algames_b200.h (the same entry points the Julia `ccall` wrapper binds).

The shipped library is `libalgames_b200.so` next to this file (built by `__graft_entry__.build()` with nvcc for
sm_100a).  There is no CPU fallback: if the library is missing, or no CUDA device is present when a handle is
created, the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_P, MAX_N, MAX_M, MAX_WALLS, MAX_CIRCLES, NSTATS, NHIST, IPC_BYTES, MAX_RANKS = 4, 48, 16, 8, 8, 10, 10, 64, 64
MODEL_IDS = {"double_integrator": 0, "unicycle": 1, "bicycle": 2, "quadrotor": 3}
SOLVER_AUTO, SOLVER_BAND = 0, 1
# per-instance status (include/algames_b200.h): 1-3 = not converged (how the last inner loop ended), 4-5 = numerical failure
CONVERGED, MAX_OUTER, LINE_SEARCH_FAILED, STALLED, SINGULAR, NONFINITE = range(6)
STATUS_NAMES = {0: "converged", 1: "max_outer", 2: "line_search_failed", 3: "stalled", 4: "singular", 5: "nonfinite"}

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libalgames_b200.so")


class ProblemDesc(C.Structure):
    _fields_ = [
        ("model", C.c_int), ("p", C.c_int), ("d", C.c_int), ("N", C.c_int),
        ("dt", C.c_double), ("lf", C.c_double), ("lr", C.c_double),
        ("Q", C.c_double * MAX_N), ("R", C.c_double * MAX_M), ("xf", C.c_double * MAX_N), ("uf", C.c_double * MAX_M),
        ("has_collision_cost", C.c_int),
        ("cc_radius", C.c_double * MAX_P), ("cc_mu", C.c_double * MAX_P),
        ("col_radius", (C.c_double * MAX_P) * MAX_P),
        ("has_control_bound", C.c_int),
        ("u_max", C.c_double * MAX_M), ("u_min", C.c_double * MAX_M),
        ("has_state_bound", C.c_int * MAX_P),
        ("x_max", (C.c_double * MAX_N) * MAX_P), ("x_min", (C.c_double * MAX_N) * MAX_P),
        ("n_walls", C.c_int * MAX_P),
        ("walls", ((C.c_double * 6) * MAX_WALLS) * MAX_P),
        ("n_circles", C.c_int * MAX_P),
        ("circles", ((C.c_double * 3) * MAX_CIRCLES) * MAX_P),
        ("x_max_con", (C.c_int * MAX_N) * MAX_P), ("x_min_con", (C.c_int * MAX_N) * MAX_P),
        ("quad_mass", C.c_double), ("spherical_collision", C.c_int),
        ("n_walls3d", C.c_int * MAX_P), ("walls3d", ((C.c_double * 12) * MAX_WALLS) * MAX_P),
        ("n_cylinders", C.c_int * MAX_P), ("cylinders", ((C.c_double * 6) * MAX_WALLS) * MAX_P),
        ("solver", C.c_int),
    ]


class OptionsC(C.Structure):
    _fields_ = [
        ("reg_0", C.c_double), ("regularize", C.c_int), ("alpha_decrease", C.c_double), ("beta", C.c_double),
        ("ls_iter", C.c_int), ("delta_min", C.c_double), ("rho_0", C.c_double), ("rho_increase", C.c_double),
        ("rho_max", C.c_double), ("lambda_max", C.c_double), ("alpha_dual", C.c_double),
        ("alphax_dual", C.c_double * MAX_P), ("active_set_tolerance", C.c_double),
        ("eps_dyn", C.c_double), ("eps_sta", C.c_double), ("eps_con", C.c_double), ("eps_opt", C.c_double),
        ("outer_iter", C.c_int), ("inner_iter", C.c_int), ("dual_reset", C.c_int),
    ]


class IBROptionsC(C.Structure):
    _fields_ = [("ibr_iter", C.c_int), ("ordering", C.c_int * MAX_P), ("delta_min", C.c_double)]


class Sizes(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("n", "m", "p", "N", "S", "nrow", "nrow_state", "nrow_control")]


class DeviceView(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in
                ("Z_dev", "L_dev", "conlam_dev", "conmu_dev", "stats_dev", "status_dev", "x0_dev", "Z0_dev", "L0_dev",
                 "results_dev")] + [("results_bytes", C.c_ulonglong)]


# every symbol include/algames_b200.h declares: name -> (restype, argtypes)
_DP, _IP, _H = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
SYMBOLS = {
    "agb_default_options": (None, [C.POINTER(OptionsC)]),
    "agb_sizes_of": (C.c_int, [C.POINTER(ProblemDesc), C.POINTER(Sizes)]),
    "agb_create": (C.c_int, [C.POINTER(ProblemDesc), C.c_int, C.c_int, C.POINTER(_H)]),
    "agb_destroy": (None, [_H]),
    "agb_last_error": (C.c_char_p, [_H]),
    "agb_get_sizes": (C.c_int, [_H, C.POINTER(Sizes)]),
    "agb_set_instance_params": (C.c_int, [_H, _DP, _DP, _DP, _DP, _DP]),
    "agb_set_initial": (C.c_int, [_H, _DP, _DP, _DP, _DP]),
    "agb_get_state": (C.c_int, [_H, _DP, _DP, _DP, _DP]),
    "agb_shift_initial": (C.c_int, [_H, C.c_int, _DP, _DP]),
    "agb_mpc_advance": (C.c_int, [_H, C.c_int, _DP, _DP, _DP]),
    "agb_mpc_advance_async": (C.c_int, [_H, C.c_int, C.c_void_p]),
    "agb_mpc_run_async": (C.c_int, [_H, C.POINTER(OptionsC), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agb_mpc_run": (C.c_int, [_H, C.POINTER(OptionsC), C.c_int, C.c_int, _DP, _DP, _IP, _DP]),
    "agb_get_stream": (C.c_void_p, [_H]),
    "agb_join_stream": (C.c_int, [_H, C.c_void_p]),
    "agb_rollout": (C.c_int, [_H]),
    "agb_residual": (C.c_int, [_H, C.c_double, C.c_double, C.c_double, _DP, _DP]),
    "agb_residual_jacobian_dense": (C.c_int, [_H, C.c_double, C.c_double, _DP]),
    "agb_kkt_solve": (C.c_int, [_H, C.c_double, C.c_double, _DP]),
    "agb_line_search": (C.c_int, [_H, C.POINTER(OptionsC), C.c_double, C.c_double, _DP, _IP]),
    "agb_update_traj": (C.c_int, [_H, _DP, _DP]),
    "agb_dual_update": (C.c_int, [_H, C.POINTER(OptionsC)]),
    "agb_penalty_update": (C.c_int, [_H, C.POINTER(OptionsC)]),
    "agb_reset_duals_penalties": (C.c_int, [_H, C.POINTER(OptionsC)]),
    "agb_evaluate_constraints": (C.c_int, [_H, _DP]),
    "agb_active_set": (C.c_int, [_H, C.c_double, C.POINTER(C.c_ubyte)]),
    "agb_active_set_sizes": (C.c_int, [_H, _IP, _IP]),
    "agb_active_set_residual": (C.c_int, [_H, _DP]),
    "agb_active_set_jacobian_dense": (C.c_int, [_H, _DP]),
    "agb_active_set_masks": (C.c_int, [_H, C.c_double, C.POINTER(C.c_ubyte), C.POINTER(C.c_ubyte)]),
    "agb_update_nullspace": (C.c_int, [_H, C.c_double, C.c_double, C.c_int, _DP, _IP]),
    "agb_band_info": (C.c_int, [_H, _IP, _IP]),
    "agb_debug_gain_solve": (C.c_int, [_H, _DP, _DP, _IP]),
    "agb_newton_solve_batch": (C.c_int, [_H, C.POINTER(OptionsC), _DP, _DP, _DP, _DP, _DP, _IP]),
    "agb_ibr_newton_solve_batch": (C.c_int, [_H, C.POINTER(OptionsC), C.POINTER(IBROptionsC), _DP, _DP, _DP, _DP, _DP, _IP]),
    "agb_ibr_residual": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, C.c_double, _DP, _DP]),
    "agb_ibr_kkt_solve": (C.c_int, [_H, C.c_int, C.c_double, C.c_double, _DP]),
    "agb_solve_from_host": (C.c_int, [_H, C.POINTER(OptionsC), _DP, _DP, _DP, _DP, _DP, _DP, _DP, _DP, _IP]),
    "agb_newton_solve_async": (C.c_int, [_H, C.POINTER(OptionsC), C.c_void_p]),
    "agb_get_device_view": (C.c_int, [_H, C.POINTER(DeviceView)]),
    "agb_set_history": (C.c_int, [_H, C.c_int]),
    "agb_get_history": (C.c_int, [_H, _DP, _IP]),
    "agb_abi_check": (C.c_int, [_IP, C.c_int]),
    "agb_abi_layout": (C.c_int, [_IP, C.c_int]),
    "agb_violations": (C.c_int, [_H, _DP, _DP, _DP, _DP]),
    "agb_peer_init": (C.c_int, [_H, C.c_int, C.c_int, _IP]),
    "agb_peer_export": (C.c_int, [_H, C.POINTER(C.c_ubyte)]),
    "agb_peer_connect": (C.c_int, [_H, C.POINTER(C.c_ubyte)]),
    "agb_peer_connect_local": (C.c_int, [C.POINTER(_H), C.c_int]),
    "agb_allgather": (C.c_int, [_H, C.c_void_p]),
    "agb_allgather_wait": (C.c_int, [_H]),
    "agb_gathered_view": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_ulonglong)]),
    "agb_unpack_gathered": (C.c_int, [_H, C.c_int, _DP, _DP, _DP, _IP]),
    "agb_create_sharded": (C.c_int, [C.POINTER(ProblemDesc), C.c_int, C.c_int, _IP, C.POINTER(_H), _IP]),
    "agb_measure_fp64_peak": (C.c_int, [C.c_int, C.c_int, _DP, C.POINTER(C.c_float)]),
    "agb_launch_count": (C.c_longlong, [_H]),
    "agb_last_solve_ms": (C.c_float, [_H]),
}

_cache = {}


def abi_layout():
    """The AGB_ABI_WORDS values agb_abi_check expects, computed from the ctypes mirrors above."""
    off = lambda st, f: getattr(st, f).offset
    return [C.sizeof(ProblemDesc), off(ProblemDesc, "dt"), off(ProblemDesc, "Q"), off(ProblemDesc, "col_radius"),
            off(ProblemDesc, "has_state_bound"), off(ProblemDesc, "walls"), off(ProblemDesc, "circles"), off(ProblemDesc, "x_max_con"),
            off(ProblemDesc, "quad_mass"), off(ProblemDesc, "solver"),
            C.sizeof(OptionsC), off(OptionsC, "alphax_dual"), off(OptionsC, "eps_dyn"), off(OptionsC, "dual_reset"),
            C.sizeof(IBROptionsC), off(IBROptionsC, "delta_min"), C.sizeof(Sizes), C.sizeof(DeviceView),
            MAX_P, MAX_N, MAX_M, MAX_WALLS, MAX_CIRCLES, NSTATS, NHIST, IPC_BYTES]


class LibraryMissing(RuntimeError):
    pass


def load(path: str | None = None) -> C.CDLL:
    """Load the C-ABI library and bind every declared symbol.  `path` defaults to the in-tree nvcc build."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise LibraryMissing(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  algames_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    # the structs above are hand-written mirrors of include/algames_b200.h: let the library compare sizes and offsets
    words = (C.c_int * len(abi_layout()))(*abi_layout())
    if lib.agb_abi_check(words, len(words)) != 0:
        raise RuntimeError(f"{path}: {lib.agb_last_error(None).decode()}")
    _cache[path] = lib
    return lib


def dptr(a):
    """double* of a C-contiguous float64 numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype.name == "float64" and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_DP)


def iptr(a):
    if a is None:
        return None
    assert a.dtype.name == "int32" and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_IP)
