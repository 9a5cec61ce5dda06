"""CPU-only: the SDF program lowering (seismicmesh_b200.geometry) and the per-point arithmetic
header shared with the CUDA kernels (csrc/dm_sdf.cuh, compiled for the host by tests/hostsim)
against the golden vectors generated from the reference."""
import ctypes as C

import numpy as np
import pytest
from conftest import load_golden, load_sdf_specs, np_ptr, relerr

from oracle import distmesh_oracle as orc
from seismicmesh_b200 import _lib, geometry

SPECS = load_sdf_specs()


def _host_sdf(hostsim, obj, x):
    prog = np.ascontiguousarray(obj.program())
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(len(x))
    hostsim.hs_sdf_eval(np_ptr(prog), np_ptr(x), C.c_long(len(x)), C.c_int(obj.dim), np_ptr(out))
    return out


@pytest.mark.parametrize("i", range(len(SPECS)))
def test_lowered_sdf_matches_reference(hostsim, i):
    g = load_golden("sdf_cases.npz")
    obj = geometry.from_spec(SPECS[i])
    d = _host_sdf(hostsim, obj, g[f"x{i}"])
    assert relerr(d, g[f"d{i}"]) < 1e-13


@pytest.mark.parametrize("i", range(len(SPECS)))
def test_bbox_and_corners_match_reference(i):
    g = load_golden("sdf_cases.npz")
    obj = geometry.from_spec(SPECS[i])
    assert np.allclose(np.asarray(obj.bbox, dtype=float), g[f"bbox{i}"], rtol=1e-14, atol=1e-14)
    c = np.zeros((0, obj.dim)) if obj.corners is None else np.asarray(obj.corners, dtype=float)
    assert c.shape == g[f"corners{i}"].shape
    assert np.allclose(c, g[f"corners{i}"], rtol=1e-14, atol=1e-14)


def test_unrotated_primitives_bit_exact(hostsim):
    g = load_golden("sdf_cases.npz")
    for i, spec in enumerate(SPECS):
        if spec[0] in ("disk", "ball", "rectangle", "cube", "torus", "prism", "cylinder"):
            prm = spec[1]
            if any(prm.get("rotate") or [0]) or prm.get("stretch") is not None:
                continue
            obj = geometry.from_spec(spec)
            assert np.array_equal(_host_sdf(hostsim, obj, g[f"x{i}"]), g[f"d{i}"]), spec[0]


def _size_struct(g):
    axes = [np.ascontiguousarray(g[k]) for k in ("axis0", "axis1", "axis2") if k in g]
    grid = np.ascontiguousarray(g["grid"])
    f = _lib.DmSizeFn()
    f.kind = _lib.SIZE_GRID
    f.dim = len(axes)
    for k, a in enumerate(axes):
        f.n[k] = len(a)
        f.axis[k] = a.ctypes.data
    f.grid = grid.ctypes.data
    return f, (axes, grid)


@pytest.mark.parametrize("name", ["interp_2d.npz", "interp_3d.npz", "r0m_values.npz"])
def test_interp_arithmetic_bit_exact(hostsim, name):
    g = load_golden(name)
    f, keep = _size_struct(g)
    x = np.ascontiguousarray(g["x"])
    out = np.empty(len(x))
    hostsim.hs_size_eval(C.byref(f), np_ptr(x), C.c_long(len(x)), np_ptr(out))
    assert np.array_equal(out, g["h"])


@pytest.mark.parametrize("dim,n,seed", [(2, 5395, 0), (2, 3, 1), (3, 676, 2), (3, 2, 3)])
def test_interp_index_search_any_axis(hostsim, dim, n, seed):
    """grid_find guesses the cell index in float32 and fixes it up against the real axis: the result
    must be SciPy's for long float32-rounded axes (BP2004 / EAGE lengths), tiny ones, points on nodes,
    between nodes one ulp either side, and far outside (linear extrapolation)."""
    from scipy.interpolate import RegularGridInterpolator

    rng = np.random.default_rng(seed)
    lo = rng.uniform(-13000.0, -10.0, dim)
    hi = lo + rng.uniform(50.0, 80000.0, dim)
    shape = [n, max(2, n // 3)] + ([max(2, n // 7)] if dim == 3 else [])
    axes = [np.linspace(lo[k], hi[k], shape[k], dtype=np.float32).astype(np.float64) for k in range(dim)]
    grid = np.ascontiguousarray(rng.uniform(20.0, 900.0, shape))
    pts = [rng.uniform(lo - 0.3 * (hi - lo), hi + 0.3 * (hi - lo), (4000, dim))]
    nodes = np.stack([a[rng.integers(0, len(a), 2000)] for a in axes], axis=1)
    pts += [nodes, np.nextafter(nodes, np.inf), np.nextafter(nodes, -np.inf)]
    pts.append(np.array([[a[0] for a in axes], [a[-1] for a in axes], [a[0] - 1e9 for a in axes], [a[-1] + 1e9 for a in axes]]))
    x = np.ascontiguousarray(np.vstack(pts))
    f = _lib.DmSizeFn()
    f.kind, f.dim = _lib.SIZE_GRID, dim
    for k, a in enumerate(axes):
        f.n[k] = len(a)
        f.axis[k] = a.ctypes.data
    f.grid = grid.ctypes.data
    out = np.empty(len(x))
    hostsim.hs_size_eval(C.byref(f), np_ptr(x), C.c_long(len(x)), np_ptr(out))
    ref = RegularGridInterpolator(tuple(axes), grid, bounds_error=False, fill_value=None)(x)
    if dim == 2:
        assert np.array_equal(out, ref)  # SciPy's compiled 2-D path: same accumulation order, bit exact
    else:
        assert np.array_equal(out, orc.interp_grid(axes, grid, x))  # the oracle restates SciPy's N-D path
        assert np.allclose(out, ref, rtol=1e-13, atol=0)


def test_projection_arithmetic(hostsim):
    g = load_golden("loop_2d.npz")
    obj = geometry.Disk([0.0, 0.0], 1.0)
    prog = np.ascontiguousarray(obj.program())
    p = np.ascontiguousarray(g["p_upd"].copy())
    h0 = float(g["h0"])
    deps = np.sqrt(np.finfo(np.double).eps) * h0
    hostsim.hs_project(np_ptr(prog), np_ptr(p), C.c_long(len(p)), C.c_int(2), C.c_double(deps), C.c_double(h0), C.c_int(0))
    assert np.array_equal(p, g["p_new"])


def test_sliver_arithmetic(hostsim):
    g = load_golden("sliver_3d.npz")
    p, t = np.ascontiguousarray(g["p"]), np.ascontiguousarray(g["t"], dtype=np.int32)
    ang = np.empty(6 * len(t))
    hostsim.hs_dihedral(np_ptr(p), np_ptr(t), C.c_long(len(t)), np_ptr(ang))
    assert relerr(ang, g["dh"]) < 1e-13
    ele = np.ascontiguousarray(g["ele"], dtype=np.int32)
    gr = np.empty((len(ele), 3))
    hostsim.hs_circumsphere_grad(np_ptr(p), np_ptr(t), np_ptr(ele), C.c_long(len(ele)), np_ptr(gr))
    assert np.array_equal(gr, g["grad"])


def test_program_depth_guard():
    d = geometry.Disk([0.0, 0.0], 1.0)
    deep = d
    for _ in range(12):  # right-nested unions need one stack slot per level
        deep = geometry.Union([geometry.Disk([0.0, 0.0], 1.0), deep])
    with pytest.raises(geometry.SDFProgramError):
        deep.program()
    assert geometry.lower(lambda x: x) is None


def test_unique_rows_matches_numpy():
    """meshutil._unique_rows (packed-key / lexsort) == np.unique(axis=0) incl. index, inverse, counts."""
    from seismicmesh_b200.meshutil import _unique_rows

    rng = np.random.default_rng(0)
    cases = [rng.integers(0, 50, (2000, 2)), rng.integers(0, 30, (3000, 3)).astype(np.int32), rng.integers(-5, 5, (500, 2)),
             np.round(rng.random((3000, 2)) * 20) / 20, rng.integers(0, 2**40, (100, 3)), np.zeros((0, 2), int),
             rng.integers(0, 9, (1, 3)), rng.integers(0, 2**30, (400, 4))]
    for a in cases:
        got = _unique_rows(a, return_index=True, return_inverse=True, return_counts=True)
        exp = np.unique(a, axis=0, return_index=True, return_inverse=True, return_counts=True)
        for x, y in zip(got, exp):
            assert np.array_equal(np.asarray(x).ravel(), np.asarray(y).ravel())
        assert got[0].dtype == exp[0].dtype and got[0].shape == exp[0].shape
        assert np.array_equal(_unique_rows(a), exp[0])
