"""ctypes binding of libdistmesh_host.so (C ABI: include/distmesh_host.h): the host side of the
retriangulation step (north_star keeps Delaunay on the host).  Loaded on first use."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DM_HOST_LIB_PATH") or os.path.join(_HERE, "libdistmesh_host.so")

_P, _I64 = C.c_void_p, C.c_int64
_SIGNATURES = {
    "dmh_version": (C.c_char_p, []),
    "dmh_delaunay2d_max_cells": (_I64, [_I64]),
    "dmh_delaunay2d": (C.c_int, [_P, _I64, _P, _I64, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "dmh_delaunay3d_max_cells": (_I64, [_I64]),
    "dmh_delaunay3d": (C.c_int, [_P, _I64, _P, _I64, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "dmh_delaunay3d_mt": (C.c_int, [_P, _I64, _P, _I64, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64), C.c_int]),
    "dmh_dt3_build": (C.c_void_p, [_P, _I64, C.c_int, C.POINTER(C.c_int)]),
    "dmh_dt3_insert": (C.c_int, [_P, _P, _I64]),
    "dmh_dt3_points": (_I64, [_P]),
    "dmh_dt3_cells": (C.c_int, [_P, _P, _I64, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "dmh_dt3_free": (None, [_P]),
    "dmh_sort_unique_rows_i32": (C.c_int, [_P, _I64, C.c_int, _I64, _P, C.POINTER(_I64), C.c_int]),
    "dmh_orient3d": (C.c_double, [_P, _P, _P, _P]),
    "dmh_insphere": (C.c_double, [_P, _P, _P, _P, _P]),
    "dmh_orient2d": (C.c_double, [_P, _P, _P]),
    "dmh_incircle": (C.c_double, [_P, _P, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it first (python -c 'import __graft_entry__ as g; g.build()' "
                "or seismicmesh_b200/csrc/host/build.sh)."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib
