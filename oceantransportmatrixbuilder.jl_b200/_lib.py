"""ctypes binding of libotmb.so (include/otmb.h).  Loading fails loudly: there is no CPU path."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libotmb.so"

OK = 0
ERR_TADV_NAN, ERR_TKH_NAN, ERR_TKVML_NAN, ERR_TKVDEEP_NAN, ERR_RHO_NAN = 1, 2, 3, 4, 5
ERR_UNKNOWN_GRID, ERR_ALL_FILL, ERR_DRY_NEIGHBOUR, ERR_BADARG, ERR_STATE = 6, 7, 8, 9, 10
ERR_COMM = 11
ERR_CUDA, ERR_NO_GPU, ERR_TOO_LARGE = 100, 101, 102
TOPO = {"bipolar": 0, "tripolar": 1, "unknown": 2}
PATH = {"fused": 0, "fused2": 1, "coo": 2}
MAT = {"T": 0, "Tadv": 1, "TκH": 2, "TκVML": 3, "TκVdeep": 4}

EXPORTS = [
    "otmb_version", "otmb_device_count", "otmb_status_string", "otmb_create", "otmb_destroy", "otmb_last_error",
    "otmb_host_alloc", "otmb_host_free", "otmb_host_trim", "otmb_set_grid", "otmb_makeindices", "otmb_get_indices",
    "otmb_gridmetrics", "otmb_set_gridmetrics", "otmb_facefluxes", "otmb_set_facefluxes", "otmb_set_mlotst",
    "otmb_set_rho3d", "otmb_transportmatrix_build", "otmb_transportmatrix_fetch", "otmb_transportmatrix_fetch_all", "otmb_host_widen", "otmb_set_operator",
    "otmb_sparse_build", "otmb_sparse_fetch", "otmb_spadd_build", "otmb_spadd_fetch", "otmb_triad_derivative",
    "otmb_dyad_derivative", "otmb_bolus_gm_velocity", "otmb_timer_start", "otmb_timer_stop", "otmb_l2_flush",
    "otmb_launch_count", "otmb_last_build_ms", "otmb_set_build_timing", "otmb_synchronize", "otmb_set_slab", "otmb_slab_counts",
    "otmb_set_rank_offset", "otmb_facefluxes_slab", "otmb_velocity2fluxes", "otmb_fluxes2velocity", "otmb_bgrid_to_cgrid", "otmb_lump_and_spray_build", "otmb_lump_and_spray_fetch", "otmb_spmv",
    "otmb_plan_slabs", "otmb_set_slab_rows", "otmb_comm_unique_id", "otmb_comm_init", "otmb_comm_free", "otmb_comm_allgather_i64", "otmb_comm_chain_transport",
    "otmb_sharded_makeindices", "otmb_set_masstransport", "otmb_sharded_facefluxes", "otmb_sharded_facefluxes_enqueue",
    "otmb_sharded_transportmatrix_build", "otmb_result_checksum",
    "otmb_facefluxes_gm", "otmb_transportmatrix_stream", "otmb_coarsen_build", "otmb_coarsen_fetch", "otmb_transportmatrix_dump",
    "otmb_selftest_division",
]


class TMParams(C.Structure):
    _fields_ = [("kH", C.c_double), ("kVML", C.c_double), ("kVdeep", C.c_double), ("rho", C.c_double),
                ("upwind", C.c_int32), ("index_base", C.c_int32), ("path", C.c_int32), ("build_mask", C.c_int32)]


_lib = None


def load():
    """Load libotmb.so.  Raises if it has not been built — the product has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                          "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, i64, dbl, i32 = C.c_void_p, C.c_int64, C.c_double, C.c_int32
    pi64 = C.POINTER(C.c_int64)
    sig = {
        "otmb_version": ([], C.c_int),
        "otmb_device_count": ([C.POINTER(C.c_int)], C.c_int),
        "otmb_status_string": ([C.c_int], C.c_char_p),
        "otmb_create": ([C.POINTER(vp), C.c_int], C.c_int),
        "otmb_destroy": ([vp], C.c_int),
        "otmb_last_error": ([vp], C.c_char_p),
        "otmb_host_alloc": ([C.POINTER(vp), i64], C.c_int),
        "otmb_host_free": ([vp], C.c_int),
        "otmb_host_trim": ([], C.c_int),
        "otmb_set_grid": ([vp, i64, i64, i64, C.c_int], C.c_int),
        "otmb_makeindices": ([vp, vp, pi64], C.c_int),
        "otmb_get_indices": ([vp, vp, vp, vp], C.c_int),
        "otmb_gridmetrics": ([vp] * 12, C.c_int),
        "otmb_set_gridmetrics": ([vp] * 9, C.c_int),
        "otmb_facefluxes": ([vp, vp, vp, dbl] + [vp] * 6, C.c_int),
        "otmb_set_facefluxes": ([vp, C.POINTER(vp)], C.c_int),
        "otmb_set_mlotst": ([vp, vp], C.c_int),
        "otmb_set_rho3d": ([vp, vp], C.c_int),
        "otmb_transportmatrix_build": ([vp, C.POINTER(TMParams), pi64], C.c_int),
        "otmb_transportmatrix_fetch": ([vp, C.c_int, vp, vp, vp], C.c_int),
        "otmb_transportmatrix_fetch_all": ([vp, C.c_int, vp, vp, vp], C.c_int),
        "otmb_host_widen": ([vp, vp, i64, i32], C.c_int),
        "otmb_set_operator": ([vp, C.c_int, i64, vp, vp, vp, i32], C.c_int),
        "otmb_sparse_build": ([vp, i64, vp, vp, vp, i64, pi64], C.c_int),
        "otmb_sparse_fetch": ([vp, vp, vp, vp], C.c_int),
        "otmb_spadd_build": ([vp, i64, vp, vp, vp, vp, vp, vp, pi64], C.c_int),
        "otmb_spadd_fetch": ([vp, vp, vp, vp], C.c_int),
        "otmb_triad_derivative": ([vp, vp, C.c_int, vp], C.c_int),
        "otmb_dyad_derivative": ([vp, vp, vp], C.c_int),
        "otmb_bolus_gm_velocity": ([vp, vp, dbl, dbl, vp, vp], C.c_int),
        "otmb_timer_start": ([vp], C.c_int),
        "otmb_timer_stop": ([vp, C.POINTER(C.c_float)], C.c_int),
        "otmb_l2_flush": ([vp], C.c_int),
        "otmb_launch_count": ([vp, pi64], C.c_int),
        "otmb_last_build_ms": ([vp, C.POINTER(C.c_float)], C.c_int),
        "otmb_set_build_timing": ([vp, i32], C.c_int),
        "otmb_synchronize": ([vp], C.c_int),
        "otmb_selftest_division": ([vp, i64, C.c_uint64, pi64, vp], C.c_int),
        "otmb_velocity2fluxes": ([vp, vp, vp, vp, dbl, vp, vp], C.c_int),
        "otmb_fluxes2velocity": ([vp, vp, vp, vp, dbl, vp, vp], C.c_int),
        "otmb_bgrid_to_cgrid": ([vp, vp, vp, dbl, vp, vp], C.c_int),
        "otmb_lump_and_spray_build": ([vp, i64, i64, i64, vp, vp, vp, i32, i32, pi64], C.c_int),
        "otmb_lump_and_spray_fetch": ([vp] * 8, C.c_int),
        "otmb_spmv": ([vp, C.c_int, C.c_int, vp, vp], C.c_int),
        "otmb_set_slab": ([vp, i64, i64], C.c_int),
        "otmb_set_slab_rows": ([vp, i64, i64], C.c_int),
        "otmb_coarsen_build": ([vp, C.c_int, pi64, pi64], C.c_int),
        "otmb_coarsen_fetch": ([vp, vp, vp, vp], C.c_int),
        "otmb_transportmatrix_dump": ([vp, C.c_int, C.c_char_p], C.c_int),
        "otmb_transportmatrix_stream": ([vp, C.POINTER(TMParams), C.POINTER(vp), vp, vp, i32, pi64, vp, vp, vp, pi64], C.c_int),
        "otmb_facefluxes_gm": ([vp, vp, vp, dbl, vp, dbl, dbl, dbl, i32] + [vp] * 8, C.c_int),
        "otmb_result_checksum": ([vp, C.c_int, i64, i64, C.POINTER(C.c_uint64)], C.c_int),
        "otmb_plan_slabs": ([vp, i64, i64, i64, i32, i32, pi64, pi64], C.c_int),
        "otmb_comm_unique_id": ([vp], C.c_int),
        "otmb_comm_init": ([vp, i32, i32, vp], C.c_int),
        "otmb_comm_free": ([vp], C.c_int),
        "otmb_comm_allgather_i64": ([vp, pi64, i32, pi64], C.c_int),
        "otmb_comm_chain_transport": ([vp, C.POINTER(i32)], C.c_int),
        "otmb_sharded_makeindices": ([vp, vp, pi64, pi64, pi64], C.c_int),
        "otmb_set_masstransport": ([vp, vp, vp, dbl], C.c_int),
        "otmb_sharded_facefluxes": ([vp, i32] + [vp] * 6, C.c_int),
        "otmb_sharded_facefluxes_enqueue": ([vp, i32], C.c_int),
        "otmb_sharded_transportmatrix_build": ([vp, C.POINTER(TMParams), pi64, pi64, pi64], C.c_int),
        "otmb_slab_counts": ([vp, pi64, pi64], C.c_int),
        "otmb_set_rank_offset": ([vp, i64], C.c_int),
        "otmb_facefluxes_slab": ([vp, vp, vp, dbl, vp, vp, i32, C.POINTER(i32)] + [vp] * 6, C.c_int),
    }
    for name, (args, res) in sig.items():
        f = getattr(L, name)          # AttributeError if the symbol is not exported
        f.argtypes = args
        f.restype = res
    _lib = L
    return L
