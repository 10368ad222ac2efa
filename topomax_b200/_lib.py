"""ctypes binding of ``libtopomax_b200.so`` (the C ABI in ``include/topomax_b200.h``).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable the
import of the engine raises, loudly, with the reason.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_int, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtopomax_b200.so")

TM_F64, TM_F32 = 0, 1
SIDE_LEFT, SIDE_RIGHT, SIDE_TOP, SIDE_BOTTOM = 1, 2, 4, 8
PRECOND_JACOBI, PRECOND_MULTIGRID = 0, 1
OPT_PRECOND, OPT_CHEB_DEGREE, OPT_CHECK_EVERY, OPT_MG_COARSE_CELLS = 1, 2, 3, 4
OPT_PROFILE = 5
OPT_P2P = 130
OPT_CYCLE_FIRST, OPT_CYCLE_LAST, OPT_CYCLE_GAMMA, OPT_FP_FLOOR_FACTOR = 133, 134, 135, 137
LEDGER_CATEGORIES = 40
OPT_CHEB_RATIO, OPT_EIG_SAFETY = 100, 101

ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOT_CONVERGED = -1, -2, -3, -4

# every symbol include/topomax_b200.h declares (tests check the library exports them all)
EXPORTED_SYMBOLS = (
    "tm_create", "tm_destroy", "tm_set_stream", "tm_set_option", "tm_last_error", "tm_version",
    "tm_load_vector", "tm_filter_apply", "tm_elast_matvec", "tm_elast_diag", "tm_state_solve",
    "tm_dot_p2", "tm_sens_rhs", "tm_md_halfstep", "tm_md_volume", "tm_md_apply", "tm_md_project", "tm_integrate", "tm_sample_field",
    "tm_last_solve_stats", "tm_mg_debug", "tm_mg_level_info", "tm_profile_read", "tm_launch_count",
    "tm_ledger_read", "tm_comm_unique_id", "tm_comm_init", "tm_local_layout", "tm_dem_strain_energy",
    "tm_fluid_create", "tm_fluid_destroy", "tm_fluid_set_stream", "tm_fluid_set_option", "tm_fluid_set_density", "tm_fluid_state_solve",
    "tm_fluid_objective", "tm_fluid_sens_rhs", "tm_fluid_apply", "tm_p2p_selftest",
)


class TmConfig(Structure):
    _fields_ = [
        ("nx", c_int), ("ny", c_int),
        ("width", c_double), ("height", c_double),
        ("lame_lambda", c_double), ("lame_mu", c_double),
        ("simp_min", c_double), ("filter_radius", c_double),
        ("fixed_sides", c_int), ("dtype", c_int), ("device", c_int),
        ("rank", c_int), ("nranks", c_int), ("mg_dist_levels", c_int),
    ]


class TmLoads(Structure):
    _fields_ = [
        ("has_force", c_int),
        ("force_center", c_double * 2), ("force_radius", c_double), ("force_value", c_double * 2),
        ("ntractions", c_int),
        ("traction_side", c_int * 8),
        ("traction_center", c_double * 8), ("traction_length", c_double * 8),
        ("traction_value", (c_double * 2) * 8),
    ]


class EngineError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libtopomax_b200 error {code}: {message}")
        self.code = code
        self.message = message


class NotConverged(EngineError):
    pass


_lib = None


def load_library() -> ctypes.CDLL:
    """Load the CUDA library or raise: the product path never degrades to a CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). topomax_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    V, D, I = c_void_p, c_double, c_int
    sigs = {
        "tm_create": ([POINTER(TmConfig), POINTER(c_void_p)], I),
        "tm_destroy": ([V], I),
        "tm_set_stream": ([V, V], I),
        "tm_set_option": ([V, I, D], I),
        "tm_last_error": ([], c_char_p),
        "tm_version": ([], c_char_p),
        "tm_load_vector": ([V, POINTER(TmLoads), V], I),
        "tm_filter_apply": ([V, I, V, V, D, I, POINTER(I), POINTER(D)], I),
        "tm_elast_matvec": ([V, V, D, V, V], I),
        "tm_elast_diag": ([V, V, D, V], I),
        "tm_state_solve": ([V, V, D, V, V, D, I, I, POINTER(I), POINTER(D)], I),
        "tm_dot_p2": ([V, V, V, POINTER(D)], I),
        "tm_sens_rhs": ([V, V, D, V, V], I),
        "tm_md_halfstep": ([V, V, V, D, V], I),
        "tm_md_volume": ([V, V, D, POINTER(D), POINTER(D)], I),
        "tm_md_apply": ([V, V, D, V, V, V, POINTER(D), POINTER(D)], I),
        "tm_md_project": ([V, V, D, D, I, POINTER(D), POINTER(I), POINTER(I)], I),
        "tm_integrate": ([V, V, POINTER(D)], I),
        "tm_sample_field": ([V, I, V, I, I, D, D, D, D, V], I),
        "tm_last_solve_stats": ([V, POINTER(D), I], I),
        "tm_comm_unique_id": ([ctypes.c_char_p], I),
        "tm_comm_init": ([V, ctypes.c_char_p], I),
        "tm_local_layout": ([V, POINTER(I), I], I),
        "tm_profile_read": ([V, POINTER(D), I], I),
        "tm_launch_count": ([], ctypes.c_longlong),
        "tm_ledger_read": ([V, POINTER(D), I, I], I),
        "tm_mg_debug": ([V, V, I, I, V, V], I),
        "tm_mg_level_info": ([V, I, POINTER(I), POINTER(I)], I),
        "tm_dem_strain_energy": ([I, I, D, D, D, D, D, D, V, V, V, V, V, POINTER(D), V], I),
        "tm_fluid_create": ([I, I, D, D, D, D, D, I, POINTER(c_void_p)], I),
        "tm_fluid_destroy": ([V], I),
        "tm_fluid_set_stream": ([V, V], I),
        "tm_fluid_set_option": ([V, I, D], I),
        "tm_fluid_set_density": ([V, V, D], I),
        "tm_fluid_state_solve": ([V, V, D, I, V, POINTER(I), POINTER(D)], I),
        "tm_fluid_objective": ([V, V, POINTER(D)], I),
        "tm_fluid_sens_rhs": ([V, V, V, V], I),
        "tm_fluid_apply": ([V, V, V, I], I),
        "tm_p2p_selftest": ([I, I, I, I, POINTER(D)], I),
    }
    for name, (argtypes, restype) in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc: int):
    if rc == 0:
        return
    msg = load_library().tm_last_error().decode("utf-8", "replace")
    if rc == ERR_NOT_CONVERGED:
        raise NotConverged(rc, msg)
    raise EngineError(rc, msg)
