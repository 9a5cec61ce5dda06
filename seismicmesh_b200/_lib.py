"""ctypes binding of libdistmesh_b200.so (C ABI: include/distmesh_b200.h).

There is NO CPU fallback: if the CUDA library is missing, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DM_LIB_PATH: developer override to time an experimental build of the same C ABI (still CUDA-only)
LIB_PATH = os.environ.get("DM_LIB_PATH") or os.path.join(_HERE, "libdistmesh_b200.so")

DM_MAX_LEVELS = 8
DM_SDF_WORDS = 24

# opcodes (include/distmesh_b200.h)
OP_DISK, OP_BALL, OP_RECT, OP_CUBE, OP_TORUS, OP_PRISM, OP_CYLINDER = 1, 2, 3, 4, 5, 6, 7
OP_UNION, OP_SUNION, OP_INTER, OP_SINTER, OP_DIFF, OP_SDIFF = 16, 17, 18, 19, 20, 21
OP_REPEAT_BEGIN, OP_REPEAT_END = 24, 25
TF_TRANSLATE, TF_ROT0, TF_ROT1, TF_ROT2, TF_STRETCH = 1, 2, 4, 8, 16
SIZE_CONST, SIZE_GRID, SIZE_EXTERNAL = 0, 1, 2


class DmSizeFn(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("dim", C.c_int32),
        ("n", C.c_int32 * 3),
        ("_pad", C.c_int32),
        ("axis", C.c_void_p * 3),
        ("grid", C.c_void_p),
        ("hconst", C.c_double),
        ("cells", C.c_void_p),
    ]


class DmPlan(C.Structure):
    _fields_ = [
        ("N", C.c_int64),
        ("T", C.c_int64),
        ("dim", C.c_int32),
        ("_pad0", C.c_int32),
        ("K", C.c_int64),
        ("keep", C.c_void_p),
        ("zero_base", C.c_void_p),
        ("zero_bytes", C.c_size_t),
        ("cnt", C.c_void_p),
        ("sync", C.c_void_p),
        ("gdone", C.c_void_p),
        ("counters", C.c_void_p),
        ("bucket", C.c_void_p),
        ("ovf_v", C.c_void_p),
        ("ovf_e", C.c_void_p),
        ("hv", C.c_void_p),
        ("adj", C.c_void_p),
        ("heap", C.c_void_p),
        ("degs", C.c_void_p),
        ("rowptr", C.c_void_p),
        ("hslot", C.c_void_p),
        ("hbar", C.c_void_p),
        ("partials", C.c_void_p),
        ("scalars", C.c_void_p),
        ("p4", C.c_void_p),
        ("esc", C.c_void_p),
        ("scan_tmp", C.c_void_p),
        ("scan_tmp_bytes", C.c_size_t),
        ("n_rows", C.c_int64),
        ("layout", C.c_int64),
    ]


_P = C.c_void_p
_I64 = C.c_int64
_D = C.c_double
_INT = C.c_int
_SZ = C.c_size_t

_SIGNATURES = {
    "dm_sdf_eval": (_INT, [_P, _P, _I64, _INT, _P, _P]),
    "dm_size_eval": (_INT, [C.POINTER(DmSizeFn), _P, _I64, _P, _P]),
    "dm_size_build_cells": (_INT, [C.POINTER(DmSizeFn), _P, _P]),
    "dm_centroids": (_INT, [_P, _P, _I64, _INT, _P, _P]),
    "dm_cull_cells": (_INT, [_P, _P, _P, _I64, _INT, _D, _P, _P]),
    "dm_compact_cells": (_INT, [_P, _P, _I64, _INT, _P, _P, _P, _SZ, _P]),
    "dm_compact_scratch_bytes": (_SZ, [_I64]),
    "dm_dihedral": (_INT, [_P, _P, _I64, _D, _D, _P, _P, _P]),
    "dm_sliver_flags": (_INT, [_P, _P, _P, _I64, _D, _D, _D, _P, _P, _P]),
    "dm_circumsphere_grad": (_INT, [_P, _P, _P, _I64, _P, _P]),
    "dm_sliver_perturb": (_INT, [_P, _I64, _P, _P, _I64, _D, _P, _P, _P]),
    "dm_cells_lead_interior": (_INT, [_P, _P, _I64, _D, _P]),
    "dm_level_set_newton": (_INT, [_P, _P, _P, _I64, _INT, _D, _P]),
    "dm_plan_bytes": (_SZ, [_I64, _I64, _INT]),
    "dm_plan_init": (_INT, [C.POINTER(DmPlan), _I64, _I64, _INT, _P, _SZ]),
    "dm_plan_set_rows": (_INT, [C.POINTER(DmPlan), _I64]),
    "dm_plan_set_layout": (_INT, [C.POINTER(DmPlan), _INT]),
    "dm_pad": (_INT, [_P, _P, _INT, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), _INT, _D, _D, _P, _P]),
    "dm_uniform_filter": (_INT, [_P, _P, _P, _I64, _I64, _I64, C.POINTER(C.c_int), _INT, _INT, _P]),
    "dm_variance_size": (_INT, [_P, _P, _I64, _INT, _D, _D, _D, _P, _P]),
    "dm_replace_below": (_INT, [_P, _I64, _D, _D, _INT, _P, _P]),
    "dm_laplacian_work_bytes": (_SZ, [_I64]),
    "dm_laplacian_smooth": (_INT, [C.POINTER(DmPlan), _P, _I64, _P, _P, _SZ, _D, _INT, C.POINTER(C.c_int), C.POINTER(C.c_double), _P]),
    "dm_stage_prep": (_INT, [C.POINTER(DmPlan), _P, _P]),
    "dm_stage_cull_chunk": (_INT, [C.POINTER(DmPlan), _P, _P, _P, _I64, _I64, _D, _INT, _P]),
    "dm_force_iteration_tail": (
        _INT,
        [C.POINTER(DmPlan), C.POINTER(_P), _INT, C.POINTER(DmSizeFn), _P, _P, _D, _D, _D, _D, _I64, _P, _P, _P],
    ),
    "dm_stage_cull_count": (_INT, [C.POINTER(DmPlan), _P, _P, _P, _D, _INT, _P]),
    "dm_stage_build_adjacency": (_INT, [C.POINTER(DmPlan), _P]),
    "dm_stage_bar_index": (_INT, [C.POINTER(DmPlan), _P]),
    "dm_bars_pairs": (_INT, [C.POINTER(DmPlan), _P, _P]),
    "dm_bar_midpoints": (_INT, [C.POINTER(DmPlan), _P, _P, _P]),
    "dm_bar_sizes": (_INT, [C.POINTER(DmPlan), C.POINTER(DmSizeFn), _P, _P]),
    "dm_stage_bar_pass": (_INT, [C.POINTER(DmPlan), _P, C.POINTER(DmSizeFn), _P]),
    "dm_stage_vertex_update": (
        _INT,
        [C.POINTER(DmPlan), _P, _P, C.POINTER(_P), _INT, C.POINTER(DmSizeFn), _D, _D, _D, _D, _I64, _P, _P, _P],
    ),
    "dm_project_points": (_INT, [_P, _P, _I64, _INT, _D, _D, _INT, _P]),
    "dm_force_iteration": (
        _INT,
        [C.POINTER(DmPlan), C.POINTER(_P), _INT, C.POINTER(DmSizeFn), _P, _P, _P, _D, _D, _D, _D, _D, _I64, _P, _P, _P],
    ),
    "dm_force_iteration_reuse": (
        _INT,
        [C.POINTER(DmPlan), C.POINTER(_P), _INT, C.POINTER(DmSizeFn), _P, _P, _D, _D, _D, _D, _I64, _P, _P, _P],
    ),
    "dm_stage_displacement": (_INT, [C.POINTER(DmPlan), _P, _P, C.POINTER(DmSizeFn), _P]),
    "dm_force_iteration_profiled": (
        _INT,
        [C.POINTER(DmPlan), C.POINTER(_P), _INT, C.POINTER(DmSizeFn), _P, _P, _P, _D, _D, _D, _D, _D, _I64, _P, _P,
         C.POINTER(C.c_float), C.c_char_p, _INT, _INT, C.POINTER(_INT)],
    ),
    "dm_size_from_velocity": (_INT, [_P, _P, _I64, _INT, _D, _D, _D, _D, _D, _D, _D, _P, _P]),
    "dm_limgrad": (_INT, [_P, _P, _I64, _I64, _I64, _D, _D, _INT, _P, C.POINTER(_INT), _P]),
    "dm_halo_push": (_INT, [_P, _P, _I64, _INT, _P, _P]),
    "dm_halo_push2": (_INT, [_P, _INT, _P, _I64, _P, _P, _P, _I64, _P, _P, C.c_uint64, _P, _P]),
    "dm_halo_wait": (_INT, [_P, _P, C.c_uint64, _P, _P]),
    "dm_halo_select": (_INT, [_P, _P, _I64, _I64, _INT, C.POINTER(_D), _INT, _INT, _P, _P]),
    "dm_scan_scratch_bytes": (_SZ, [_I64]),
    "dm_exclusive_scan_i32": (_INT, [_P, _P, _I64, _P, _SZ, _P]),
    "dm_version": (C.c_char_p, []),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found. seismicmesh_b200 has no CPU fallback: build the sm_100a CUDA "
            "library first (python -c 'import __graft_entry__ as g; g.build()' or "
            "seismicmesh_b200/csrc/build.sh)."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class DistmeshCudaError(RuntimeError):
    pass


_ERR = {-1: "DM_ERR_ARG (bad argument)", -2: "DM_ERR_WORKSPACE (workspace too small)", -3: "DM_ERR_PROGRAM"}


def check(rc, what=""):
    if rc != 0:
        msg = _ERR.get(rc, f"cudaError {rc}")
        raise DistmeshCudaError(f"{what}: {msg}")
