"""Pins oracle/distmesh_oracle.py against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
from conftest import load_golden, load_sdf_specs, relerr

from oracle import distmesh_oracle as orc

SPECS = load_sdf_specs()


@pytest.mark.parametrize("i", range(len(SPECS)))
def test_sdf_matches_reference(i):
    g = load_golden("sdf_cases.npz")
    d = orc.sdf(SPECS[i], g[f"x{i}"])
    # un-rotated trees are bit-identical; rotated/stretched ones differ by BLAS rounding only
    assert relerr(d, g[f"d{i}"]) < 1e-13


@pytest.mark.parametrize("name", ["interp_2d.npz", "interp_3d.npz", "r0m_values.npz"])
def test_interp_bit_exact(name):
    g = load_golden(name)
    axes = [g[k] for k in ("axis0", "axis1", "axis2") if k in g]
    h = orc.interp_grid(axes, g["grid"], g["x"])
    assert np.array_equal(h, g["h"])


def test_r0m_pinned_values():
    g = load_golden("r0m_values.npz")
    h = orc.interp_grid([g["axis0"], g["axis1"]], g["grid"], g["x"])
    assert h[0] == 100 and h[1] == 150  # reference tests/test_2dmesher_r0m_values.py:42-43


def _fd_fh(name, g):
    if name == "loop_2d.npz":
        spec = ("disk", dict(x0=[0.0, 0.0], r=1.0))
        levels = [spec]
    elif name == "loop_2d_levels.npz":
        spec = ("rectangle", dict(bbox=(0.0, 1.0, 0.0, 1.0)))
        levels = [spec, ("disk", dict(x0=[0.5, 0.5], r=0.25))]
    elif name == "loop_3d.npz":
        spec = ("ball", dict(x0=[0.0, 0.0, 0.0], r=1.0))
        levels = [spec]
    elif name == "loop_2d_grid.npz":
        spec = ("rectangle", dict(bbox=tuple(g["bbox"])))
        levels = [spec]
    else:
        spec = ("cube", dict(bbox=tuple(g["bbox"])))
        levels = [spec]
    if "grid" in g:
        axes = [g[k] for k in ("axis0", "axis1", "axis2") if k in g]
        fh = lambda x: orc.interp_grid(axes, g["grid"], x)  # noqa: E731
    else:
        h0 = float(g["h0"])
        fh = lambda x: np.array([h0] * len(x))  # noqa: E731
    return [(lambda x, s=s: orc.sdf(s, x)) for s in levels], fh


LOOPS = ["loop_2d.npz", "loop_2d_levels.npz", "loop_3d.npz", "loop_2d_grid.npz", "loop_3d_grid.npz"]


@pytest.mark.parametrize("name", LOOPS)
def test_loop_body(name):
    g = load_golden(name)
    levels, fh = _fd_fh(name, g)
    p, t, h0 = g["p"], g["t"], float(g["h0"])
    geps = 0.1 * h0
    deps = np.sqrt(np.finfo(np.double).eps) * h0
    out = orc.force_iteration(p, t, levels, fh, h0, geps, deps)
    assert np.array_equal(t[out["keep"]], g["t_kept"])          # cull: bit exact
    assert np.array_equal(out["bars"], g["bars"])                # bars: bit exact
    assert out["bars"].dtype == np.int32
    assert np.array_equal(out["h"], g["hbars"])
    assert relerr(out["Ftot"], g["Ftot"]) < 1e-12
    assert relerr(out["p"], g["p_new"]) < 1e-12
    assert abs(out["maxdp"] - float(g["maxdp"])) <= 1e-12 * max(1.0, float(g["maxdp"]))


def test_scatter_order_is_bit_exact():
    """The sequential COO accumulation order (generation/utils.py:48-68) is restated exactly."""
    g = load_golden("loop_2d.npz")
    p, tk = g["p"], g["t_kept"]
    F = orc.compute_forces(p, tk, lambda x: np.array([float(g["h0"])] * len(x)), 1.2)
    assert np.array_equal(F, g["Ftot"])


def test_sliver_blocks():
    g = load_golden("sliver_3d.npz")
    p, t = g["p"], g["t"]
    dh = orc.dihedral_angles(p, t)
    assert relerr(dh, g["dh"]) < 1e-12
    ele = orc.sliver_cells(p, t, float(g["min_dh"]), float(g["max_dh"]))
    assert np.array_equal(ele, g["ele"])
    s = t[ele]
    gr = orc.circumsphere_grad(p[s[:, 0]], p[s[:, 1]], p[s[:, 2]], p[s[:, 3]])
    assert np.allclose(gr, g["grad"], rtol=1e-10, atol=0)
    pn = orc.sliver_perturbation(p, t, ele, float(g["step"]), float(g["h0"]))
    assert relerr(pn, g["p_new"]) < 1e-12


def test_initial_points():
    g = load_golden("init_points.npz")
    disk = lambda x: orc.sdf(("disk", dict(x0=[0.0, 0.0], r=1.0)), x)  # noqa: E731
    p = orc.initial_points(0.05, 0.005, 2, np.array([[-1.0, 1.0], [-1.0, 1.0]]), lambda x: np.array([0.05] * len(x)),
                           disk, np.empty((0, 2)))
    assert np.array_equal(p, g["disk"])
    ball = lambda x: orc.sdf(("ball", dict(x0=[0.0, 0.0, 0.0], r=1.0)), x)  # noqa: E731
    p = orc.initial_points(0.2, 0.02, 3, np.array([[-1.0, 1.0]] * 3), lambda x: np.array([0.2] * len(x)),
                           ball, np.empty((0, 3)))
    assert np.array_equal(p, g["ball"])


# --------------------------------------------------------------------------------------------
# sizing preprocessing: the reference's get_sizing_function_from_segy (incl. its native FastHJ
# gradient limiter) on the velocity models of tests/golden/synth.py
# --------------------------------------------------------------------------------------------
import os  # noqa: E402
import sys  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from synth import sizing_cases  # noqa: E402

SIZING = sizing_cases()


@pytest.mark.parametrize("name", sorted(SIZING))
def test_sizing_matches_reference(name):
    vp, bbox, kw = SIZING[name]
    g = load_golden(f"sizing_{name}.npz")
    cell, bb = orc.sizing_from_velocity(vp, bbox, **kw)
    assert cell.shape == g["grid"].shape and np.allclose(bb, g["bbox"], rtol=0, atol=0)
    assert cell.min() == g["gmin"] and cell.max() == g["gmax"]
    # the golden grid is stored as float32; its float64 checksum pins the rest
    assert np.abs(cell - g["grid"]).max() <= 6e-8 * g["gmax"]
    assert abs(cell.sum() - g["checksum"]) <= 1e-12 * abs(g["checksum"])


def test_limgrad_is_a_fixed_point_and_minimal():
    rng = np.random.default_rng(3)
    f0 = rng.uniform(50.0, 900.0, (23, 31, 17))
    delta, ftol = 12.5, 50.0 * np.sqrt(1e-9)
    f = orc.limgrad(f0, delta, ftol)
    assert (f <= f0).all()
    for ax in range(3):
        assert np.abs(np.diff(f, axis=ax)).max() <= delta + ftol
    # minimality: a limited value is either untouched or sits exactly delta above a neighbour
    lowered = f < f0
    m = np.full_like(f, np.inf)
    for ax in range(3):
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[ax], hi[ax] = slice(0, -1), slice(1, None)
        m[tuple(hi)] = np.minimum(m[tuple(hi)], f[tuple(lo)])
        m[tuple(lo)] = np.minimum(m[tuple(lo)], f[tuple(hi)])
    assert np.abs(f[lowered] - (m[lowered] + delta)).max() <= 2 * ftol


@pytest.mark.parametrize("shape", [(40, 55, 1), (17, 23, 12)])
def test_limgrad_oracle_vs_reference_native(shape):
    """The reference's own compiled FastHJ limiter (oracle/_ref, built from its sources) against the
    oracle's fixed-point restatement on a rough random field."""
    from oracle import ref_harness

    try:
        hj = ref_harness.load_native_fasthj()
    except ImportError:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(shape[0])
    f0 = rng.uniform(40.0, 1500.0, shape)
    elen, grade = 12.5, 0.2
    # the reference flattens in Fortran order and passes dims (sz0, sz1, sz2) (mesh_size_function.py:486-494)
    out = np.asarray(hj.limgrad([*shape], elen, grade, 10000, f0.flatten("F"))).reshape(shape, order="F")
    mine = orc.limgrad(f0, elen * grade, f0.min() * np.sqrt(1e-9))
    assert np.abs(mine - out).max() <= 4 * f0.min() * np.sqrt(1e-9)


def test_pad_restatement_equals_numpy():
    """The staged restatement of np.pad that the CUDA kernel dm_pad follows (oracle.pad_staged) against
    np.pad itself: the reference's three pad styles, its one-sided padding of axis 0, end values equal to the
    array maximum (the reference's choice) with the maximum inside and on the boundary (NumPy's zero-step rule)."""
    rng = np.random.default_rng(0)
    for shape, padding in (((7, 9), ((3, 0), (2, 2))), ((5, 6, 4), ((2, 0), (3, 3), (1, 1))), ((6, 5), ((0, 0), (4, 4))),
                           ((4, 3, 5), ((1, 0), (0, 0), (2, 2)))):
        for on_boundary in (False, True):
            a = rng.uniform(1.0, 5.0, shape)
            if on_boundary:
                a[(0,) * len(shape)] = 9.0
            ev = [float(a.max())] * 2
            assert np.array_equal(orc.pad_staged(a, padding, "edge", ev), np.pad(a, padding, "edge"))
            assert np.array_equal(orc.pad_staged(a, padding, "constant", ev), np.pad(a, padding, "constant", constant_values=tuple(ev)))
            assert np.array_equal(orc.pad_staged(a, padding, "linear_ramp", ev), np.pad(a, padding, "linear_ramp", end_values=tuple(ev)))
            assert np.array_equal(orc.pad_staged(a, padding, "linear_ramp", [7.5, 0.25]),
                                  np.pad(a, padding, "linear_ramp", end_values=(7.5, 0.25)))


def test_uniform_filter_restatement_equals_scipy():
    """oracle.uniform_filter (the recurrence dm_uniform_filter follows) against scipy.ndimage.uniform_filter,
    bit for bit: the `grad=` option's windowed mean of vp and of vp**2 (mesh_size_function.py:428-448)."""
    from scipy import ndimage

    rng = np.random.default_rng(1)
    for shape, size in (((40, 57), (10, 10)), ((23, 31, 19), (10, 10, 10)), ((15, 40), (3, 7)), ((12, 13, 14), (5, 4, 9))):
        a = rng.uniform(1500.0, 4500.0, shape)
        assert np.array_equal(orc.uniform_filter(a, size), ndimage.uniform_filter(a, size))
        assert np.array_equal(orc.uniform_filter(a**2, size), ndimage.uniform_filter(a**2, size))
