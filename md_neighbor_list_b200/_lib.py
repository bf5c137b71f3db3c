"""ctypes binding of libnlist_b200.so (include/nlist_b200.h).

There is no CPU fallback: if the CUDA library has not been built this module raises at import of the symbols, and
every compute entry point fails with NLB200_ERR_CUDA when no device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NLB200_LIB: an alternative build of the same library (tuning experiments, tools/gpu_tune.sh)
LIB_PATH = os.environ.get("NLB200_LIB") or os.path.join(HERE, "lib", "libnlist_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_OUT_OF_BOX, ERR_ELL_ROWS, ERR_STATE, ERR_CELL_CAPACITY = range(8)
F32, F64 = 0, 1
HALF_CSR, FULL_CSR, FULL_ELL_TRANSPOSED = 0, 1, 2
OPT_POSITION_STRIDE, OPT_SORT_ROWS, OPT_ELL_ROWS, OPT_EXACT_ONLY, OPT_USE_GRAPH, OPT_KERNEL_VARIANT = 1, 2, 3, 4, 5, 6
OPT_PROFILE = 7
OPT_MAX_IN_CELL = 8
OPT_PDL = 9


class Stats(C.Structure):
    _fields_ = [("n", C.c_int64), ("number_of_pairs", C.c_int64), ("candidates_tested", C.c_int64),
                ("band_tests", C.c_int64), ("required_entries", C.c_int64), ("capacity_entries", C.c_int64),
                ("mesh", C.c_int32 * 3), ("max_partners", C.c_int32), ("max_in_cell", C.c_int32),
                ("reserved", C.c_int32)]


# every symbol include/nlist_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _i32, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
SYMBOLS = {
    "nlb200_version": (C.c_int, []),
    "nlb200_status_string": (C.c_char_p, [C.c_int]),
    "nlb200_create": (C.c_int, [_dbl, _dbl, _dbl, _dbl, _i32, _i32, C.POINTER(_vp)]),
    "nlb200_set_option": (C.c_int, [_vp, _i32, _i64]),
    "nlb200_set_cell_window": (C.c_int, [_vp, _i32, C.c_int32, C.c_int32]),
    "nlb200_initialize": (C.c_int, [_vp, _i64, _i64]),
    "nlb200_reserve": (C.c_int, [_vp, _i64]),
    "nlb200_reserve_cell_capacity": (C.c_int, [_vp, _i64]),
    "nlb200_destroy": (C.c_int, [_vp]),
    "nlb200_build": (C.c_int, [_vp, _vp, _i64, _vp]),
    "nlb200_build_subset": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "nlb200_mark_enqueued": (C.c_int, [_vp, _vp]),
    "nlb200_synchronize": (C.c_int, [_vp]),
    "nlb200_build_host": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, C.POINTER(_i64)]),
    "nlb200_fetch_partners_host": (C.c_int, [_vp, _vp, _i64]),
    "nlb200_number_of_partners": (_vp, [_vp]),
    "nlb200_offsets": (_vp, [_vp]),
    "nlb200_offsets32": (_vp, [_vp]),
    "nlb200_partners": (_vp, [_vp]),
    "nlb200_ell_transposed": (_vp, [_vp]),
    "nlb200_number_of_pairs": (_i64, [_vp]),
    "nlb200_cell_start": (_vp, [_vp]),
    "nlb200_sorted_ids": (_vp, [_vp]),
    "nlb200_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "nlb200_get_stage_times": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "nlb200_stage_name": (C.c_char_p, [C.c_int]),
    "nlb200_required_entries": (_i64, [_vp]),
    "nlb200_last_error": (C.c_char_p, [_vp]),
    "nlb200_track_reference": (C.c_int, [_vp, _vp, _i64, _vp]),
    "nlb200_max_displacement": (C.c_int, [_vp, _vp, _i64, _vp, C.POINTER(_dbl)]),
    "nlb200_lj_forces": (C.c_int, [_vp, _vp, _dbl, _dbl, _dbl, _vp, _vp, _vp]),
    "nlb200_gather_sorted": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "nlb200_select_slab": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _dbl, _dbl, _vp, _i64, _vp, _vp, _i64, _vp]),
    "nlb200_pack_slab": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _i32, _dbl, _dbl, _vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "nlb200_pack_slab2": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _dbl, _dbl, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64,
                                    _vp]),
    "nlb200_pack_faces": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _dbl, _dbl, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "nlb200_p2p_alloc": (C.c_int, [_i64, C.POINTER(_vp), _vp]),
    "nlb200_p2p_open": (C.c_int, [_vp, C.POINTER(_vp)]),
    "nlb200_p2p_close": (C.c_int, [_vp]),
    "nlb200_p2p_free": (C.c_int, [_vp]),
    "nlb200_pack_faces_p2p": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _dbl, _dbl, _vp, _vp, _vp, _vp, _i64, _vp, _vp,
                                        _vp, _vp, _vp, _vp]),
    "nlb200_halo_wait": (C.c_int, [_vp, _i32, _vp]),
    "nlb200_set_halo_sync": (C.c_int, [_vp, _vp, _vp, _vp]),
    "nlb200_halo_done": (C.c_int, [_vp, _vp, _vp, _vp]),
    "nlb200_set_halo_pack": (C.c_int, [_vp, _i32, _dbl, _dbl, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nlb200_halo_refresh": (C.c_int, [_vp, _vp, _vp]),
    "nlb200_select_slab_workspace": (_i64, [_i64]),
    "nlb200_shift_axis": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _dbl, _vp]),
    "nlb200_gather_records": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "nlb200_workload_fcc": (_i64, [_dbl, _dbl, _i32, _i32, _i32, C.c_uint32, _vp, _i32, _i64]),
    "nlb200_workload_uniform": (_i64, [_i64, _dbl, C.c_uint64, _vp, _i32]),
    "nlb200_workload_clustered": (_i64, [_i64, _dbl, _i32, C.c_uint64, _vp, _i32]),
}

_lib = None


class NlistError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"nlist_b200 status {status}: {message}")
        self.status = status


def lib() -> C.CDLL:
    """Load libnlist_b200.so; fail loudly if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA library first "
                "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(handle, status: int) -> None:
    if status != OK:
        L = lib()
        msg = L.nlb200_last_error(handle).decode() if handle else ""
        raise NlistError(status, msg or L.nlb200_status_string(status).decode())
