"""Generate the committed golden vectors under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference and `make -C oracle`):

    python tests/golden/make_golden.py

Everything written here is produced by calling the *unmodified* reference functions
(imported through oracle/ref_harness.py) on fixed-seed inputs:

  sdf_cases.json/.npz   : `.eval` of every primitive x transform x combinator
                          (geometry/signed_distance_functions.py)
  interp_{2d,3d}.npz    : `SizeFunction.eval` of gridded sizing functions built by the
                          reference's `_build_sizing_function` (scipy RGI, float32 axes)
  r0m_values.npz        : the sizing grid of tests/test_2dmesher_r0m_values.py whose two
                          eval values the reference test pins to exactly 100 and 150
  loop_{2d,3d,2d_grid,3d_grid}.npz : one loop body on a Qhull mesh of jittered points:
                          `_remove_triangles_outside`, `_get_edges`(unique_edges),
                          `_compute_forces`, `p += dt*F`, `_project_points_back_newton`
  sliver_3d.npz         : `calc_dihedral_angles`, `_calc_dihedral_angles`,
                          `calc_circumsphere_grad` + the perturbation step
  init_points.npz       : `_generate_initial_points` for Disk / Ball / gridded Rectangle
  sizing_<case>.npz     : grids built by `get_sizing_function_from_segy` (incl. the native FastHJ
                          gradient limiter) for the velocity models of tests/golden/synth.py
  e2e.json              : aggregate outcomes of full `generate_mesh` / `sliver_removal` runs
                          (vertex count, cell count, min/mean quality, area) with Qhull as
                          the triangulator.
  segy_testing.npz,
  segy_tests.json       : the velocity model of the reference's SEG-Y fixture (tests/testing.segy,
                          decoded by the harness's segyio stand-in) and the outcome of the reference
                          tests that mesh it, next to the answers those tests assert  [`segy`]
  meshutil_{2d,3d}.npz  : geometry/utils.py at termination: simp_vol, simp_qual, boundary edges /
                          facets / vertices / entities, fix_mesh, delete_boundary_entities,
                          laplacian2_fixed_point  [`meshutil`]
  reference_tests.json  : the reference's known-answer tests (2dmesher_SDF, immersion, smooth_sets,
                          pfix, verbose) replayed, next to the answers they assert  [`reftests`]

`python tests/golden/make_golden.py segy|meshutil|reftests|reftests3d` regenerates only those.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402

sm = ref_harness.load_reference()
mg = sm.generation.mesh_generator
from scipy.spatial import Delaunay  # noqa: E402


def _obj_from_spec(spec):
    """Build the REFERENCE geometry object for an oracle-style spec."""
    kind = spec[0]
    if kind in ("union", "intersection", "difference"):
        cls = dict(union=sm.Union, intersection=sm.Intersection, difference=sm.Difference)[kind]
        return cls([_obj_from_spec(c) for c in spec[1]], smoothness=spec[2])
    if kind == "repeat":
        return sm.Repeat(tuple(spec[1]), _obj_from_spec(spec[2]), list(spec[3]))
    prm = dict(spec[1])
    kw = dict(
        rotate=list(prm.pop("rotate", [0.0, 0.0, 0.0])),
        stretch=None if prm.get("stretch") is None else np.array(prm["stretch"], dtype=float),
        translate=None if prm.get("translate") is None else np.array(prm["translate"], dtype=float),
    )
    prm.pop("stretch", None)
    prm.pop("translate", None)
    if kind == "disk":
        return sm.Disk(prm["x0"], prm["r"], **kw)
    if kind == "ball":
        return sm.Ball(prm["x0"], prm["r"], **kw)
    if kind == "rectangle":
        return sm.Rectangle(tuple(prm["bbox"]), **kw)
    if kind == "cube":
        return sm.Cube(tuple(prm["bbox"]), **kw)
    if kind == "torus":
        return sm.Torus(prm["r1"], prm["r2"], **kw)
    if kind == "prism":
        return sm.Prism(prm["b"], prm["h"], **kw)
    if kind == "cylinder":
        return sm.Cylinder(h=prm["h"], r=prm["r"], **kw)
    raise ValueError(kind)


def sdf_specs():
    T2 = dict(rotate=[0.3, 0.0, 0.0], stretch=[1.5, 0.5], translate=[0.2, -0.1])
    T3 = dict(rotate=[0.3, -0.2, 0.5], stretch=[1.0, 2.0, 0.5], translate=[0.1, 0.2, -0.3])
    R3 = dict(rotate=[0.0, 0.7, 0.0])
    specs = [
        ("disk", dict(x0=[0.0, 0.0], r=1.0)),
        ("disk", dict(x0=[0.2, -0.3], r=0.7, **T2)),
        ("rectangle", dict(bbox=(-1.0, 1.0, -0.5, 0.75))),
        ("rectangle", dict(bbox=(-1.0, 1.0, -0.5, 0.75), **T2)),
        ("ball", dict(x0=[0.0, 0.0, 0.0], r=1.0)),
        ("ball", dict(x0=[0.1, 0.2, 0.3], r=0.8, **T3)),
        ("cube", dict(bbox=(-1.0, 1.0, -0.5, 0.5, 0.0, 2.0))),
        ("cube", dict(bbox=(-1.0, 1.0, -0.5, 0.5, 0.0, 2.0), **T3)),
        ("torus", dict(r1=1.0, r2=0.3)),
        ("torus", dict(r1=1.0, r2=0.3, **T3)),
        ("prism", dict(b=0.6, h=0.8)),
        ("prism", dict(b=0.6, h=0.8, **R3)),
        ("cylinder", dict(h=1.5, r=0.4)),
        ("cylinder", dict(h=1.5, r=0.4, **T3)),
        ("union", [("disk", dict(x0=[0.0, 0.0], r=0.6)), ("rectangle", dict(bbox=(0.0, 1.0, 0.0, 1.0)))], 0.0),
        ("union", [("disk", dict(x0=[0.0, 0.0], r=0.6)), ("rectangle", dict(bbox=(0.0, 1.0, 0.0, 1.0))),
                   ("disk", dict(x0=[1.0, 1.0], r=0.3))], 0.2),
        ("intersection", [("disk", dict(x0=[0.0, 0.0], r=1.0)), ("rectangle", dict(bbox=(0.0, 1.0, 0.0, 1.0)))], 0.0),
        ("intersection", [("ball", dict(x0=[0.0, 0.0, 0.0], r=1.0)), ("cube", dict(bbox=(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)))], 0.1),
        ("difference", [("rectangle", dict(bbox=(0.0, 1.0, 0.0, 1.0))), ("disk", dict(x0=[0.0, 0.0], r=0.5)),
                        ("disk", dict(x0=[1.0, 1.0], r=0.25))], 0.0),
        ("difference", [("cube", dict(bbox=(-1.0, 1.0, -1.0, 1.0, -1.0, 1.0))), ("ball", dict(x0=[0.0, 0.0, 0.0], r=0.7)),
                        ("cylinder", dict(h=3.0, r=0.2))], 0.15),
        # quarter annulus of tests/test_2d_min_qual.py
        ("intersection", [("rectangle", dict(bbox=(0.0, 1.0, 0.0, 1.0))),
                          ("difference", [("disk", dict(x0=[0.0, 0.0], r=1.0)), ("disk", dict(x0=[0.0, 0.0], r=0.5))], 0.0)], 0.0),
        ("repeat", (-2.0, 2.0, -2.0, 2.0, -2.0, 2.0), ("ball", dict(x0=[0.0, 0.0, 0.0], r=0.3)), [1.0, 1.0, 1.0]),
        ("union", [("cube", dict(bbox=(0.0, 1.0, 0.0, 1.0, 0.0, 1.0))),
                   ("intersection", [("ball", dict(x0=[1.0, 1.0, 1.0], r=0.8, **R3)), ("torus", dict(r1=1.0, r2=0.4))], 0.0)], 0.05),
    ]
    return specs


def _spec_dim(spec):
    k = spec[0]
    if k in ("disk", "rectangle"):
        return 2
    if k in ("union", "intersection", "difference"):
        return _spec_dim(spec[1][0])
    return 3


def gen_sdf():
    rng = np.random.default_rng(1234)
    specs = sdf_specs()
    arrays = {}
    for i, spec in enumerate(specs):
        dim = _spec_dim(spec)
        obj = _obj_from_spec(spec)
        x = rng.uniform(-2.2, 2.2, size=(257, dim))
        arrays[f"x{i}"] = x
        arrays[f"d{i}"] = np.asarray(obj.eval(x.copy()), dtype=np.float64)
        arrays[f"bbox{i}"] = np.asarray(obj.bbox, dtype=np.float64)
        c = obj.corners
        arrays[f"corners{i}"] = np.zeros((0, dim)) if c is None else np.asarray(c, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "sdf_cases.npz"), **arrays)
    with open(os.path.join(HERE, "sdf_cases.json"), "w") as f:
        json.dump(specs, f, indent=1)


sys.path.insert(0, HERE)
from synth import sizing_cases, synth_vp_2d, synth_vp_3d  # noqa: E402


def _rgi_parts(ef):
    rgi = ef.cell_size
    return [np.asarray(g, dtype=np.float64) for g in rgi.grid], np.asarray(rgi.values, dtype=np.float64)


def size_fn_2d():
    bbox = (-3000.0, 0.0, 0.0, 8000.0)
    vp = synth_vp_2d(61, 161, bbox)
    ef = sm.get_sizing_function_from_segy(
        None, bbox, velocity_data=vp, hmin=50.0, wl=10, freq=2.0, grade=0.15, dt=0.001,
        domain_pad=500.0, pad_style="edge", nz=61, nx=161,
    )
    return ef


def size_fn_3d():
    bbox = (-2000.0, 0.0, 0.0, 4000.0, 0.0, 3000.0)
    vp = synth_vp_3d(21, 41, 31, bbox)
    ef = sm.get_sizing_function_from_segy(
        None, bbox, velocity_data=vp, hmin=150.0, wl=5, freq=2.0, grade=0.15, hmax=5e3,
        domain_pad=250.0, pad_style="linear_ramp", nz=21, nx=41, ny=31,
    )
    return ef


def gen_interp():
    rng = np.random.default_rng(99)
    for name, ef in (("interp_2d", size_fn_2d()), ("interp_3d", size_fn_3d())):
        axes, grid = _rgi_parts(ef)
        dim = len(axes)
        bb = np.array(ef.bbox).reshape(-1, 2)
        x = rng.uniform(bb[:, 0] - 300, bb[:, 1] + 300, size=(2000, dim))  # incl. out of range
        # exact node hits and cell-edge hits
        for k in range(60):
            x[k] = [axes[d][rng.integers(0, len(axes[d]))] for d in range(dim)]
        h = np.asarray(ef.eval(x), dtype=np.float64)
        out = dict(x=x, h=h, grid=grid, bbox=np.asarray(ef.bbox), hmin=ef.hmin)
        for d in range(dim):
            out[f"axis{d}"] = axes[d]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)

    # tests/test_2dmesher_r0m_values.py:15-43
    bbox = (-10000.0, 0.0, 0.0, 10000.0)
    vs = np.zeros((200, 200))
    vs[0:150, :] = 1000
    ef = sm.get_sizing_function_from_segy(
        None, bbox=bbox, grade=0.0, grad=0.0, wl=5, freq=2.0, hmin=75, hmax=10e6,
        velocity_data=vs, nz=200, nx=200,
    )
    axes, grid = _rgi_parts(ef)
    xq = np.array([[-5000.0, 5000.0], [-1.0, 5000.0]])
    hq = np.array([float(np.ravel(ef.eval((-5000, 5000)))[0]), float(np.ravel(ef.eval((-1, 5000)))[0])])
    assert hq[0] == 100 and hq[1] == 150
    np.savez_compressed(
        os.path.join(HERE, "r0m_values.npz"), axis0=axes[0], axis1=axes[1], grid=grid, x=xq, h=hq,
        bbox=np.asarray(ef.bbox),
    )


def gen_sizing():
    """sizing_<case>.npz: the gridded size function the REFERENCE's get_sizing_function_from_segy
    builds (wavelength / gradient sizing, clamps, CFL, FastHJ gradation, domain pad) for the cases of
    tests/golden/synth.py -- stored as float32 + the exact min/max (grids are smooth; 1e-7 relative
    rounding is far below the gradation tolerance the tests allow)."""
    import warnings

    for name, (vp, bbox, kw) in sizing_cases().items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ef = sm.get_sizing_function_from_segy(None, bbox, velocity_data=vp.copy(), **kw)
        axes, grid = _rgi_parts(ef)
        np.savez_compressed(
            os.path.join(HERE, f"sizing_{name}.npz"), grid=grid.astype(np.float32), gmin=grid.min(), gmax=grid.max(),
            bbox=np.asarray(ef.bbox), hmin=ef.hmin, checksum=float(grid.sum()),
        )
        print("sizing", name, grid.shape, grid.min(), grid.max())


def _jittered_mesh(fd_obj, h0, bbox, seed, dim):
    rng = np.random.default_rng(seed)
    p = sm.generation.utils.create_staggered_grid(h0, dim, np.array(bbox).reshape(-1, 2))
    p = p[fd_obj.eval(p) < 0.1 * h0]
    p = p + rng.uniform(-0.2 * h0, 0.2 * h0, size=p.shape)
    t = Delaunay(p).simplices.astype(np.int32)
    return np.ascontiguousarray(p), np.ascontiguousarray(t)


def _loop_body(p, t, fd, fh, h0, dim, levels=None):
    """The reference loop body, stage by stage, with its own functions."""
    L0mult = 1 + 0.4 / 2 ** (dim - 1)
    geps = 0.1 * h0
    deps = np.sqrt(np.finfo(np.double).eps) * h0
    tk = mg._remove_triangles_outside(p, t, fd, geps)
    bars = mg._get_edges(tk)
    Ftot = mg._compute_forces(p, tk, fh, h0, L0mult)
    pn = p + 0.30 * Ftot
    lv = [fd] if levels is None else levels
    pp = pn.copy()
    for idx, level in enumerate(lv):
        pp = mg._project_points_back_newton(pp, level, deps, h0, idx)
    maxdp = 0.30 * np.sqrt((Ftot**2).sum(1)).max()
    hbars = fh(p[bars].sum(1) / 2)
    return dict(p=p, t=t, t_kept=tk, bars=bars, Ftot=Ftot, p_upd=pn, p_new=pp, maxdp=maxdp,
                hbars=np.asarray(hbars, dtype=np.float64), h0=h0)


def gen_loop():
    # 2D disk, scalar h
    disk = sm.Disk([0.0, 0.0], 1.0)
    h0 = 0.08
    p, t = _jittered_mesh(disk, h0, disk.bbox, 5, 2)
    fh, _, _, _ = mg._unpack_sizing(h0)
    np.savez_compressed(os.path.join(HERE, "loop_2d.npz"), **_loop_body(p, t, disk.eval, fh, h0, 2))

    # 2D with an immersed subdomain level (tests/test_immersion.py)
    rect = sm.Rectangle((0.0, 1.0, 0.0, 1.0))
    sub = sm.Disk([0.5, 0.5], 0.25)
    h0 = 0.05
    p, t = _jittered_mesh(rect, h0, rect.bbox, 6, 2)
    fh, _, _, _ = mg._unpack_sizing(h0)
    np.savez_compressed(
        os.path.join(HERE, "loop_2d_levels.npz"),
        **_loop_body(p, t, rect.eval, fh, h0, 2, levels=[rect.eval, sub.eval]),
    )

    # 3D ball, scalar h
    ball = sm.Ball([0.0, 0.0, 0.0], 1.0)
    h0 = 0.25
    p, t = _jittered_mesh(ball, h0, ball.bbox, 7, 3)
    fh, _, _, _ = mg._unpack_sizing(h0)
    np.savez_compressed(os.path.join(HERE, "loop_3d.npz"), **_loop_body(p, t, ball.eval, fh, h0, 3))

    # 2D gridded fh on the padded rectangle
    ef = size_fn_2d()
    rect = sm.Rectangle(ef.bbox)
    h0 = float(ef.hmin) * 4
    p, t = _jittered_mesh(rect, h0, ef.bbox, 8, 2)
    axes, grid = _rgi_parts(ef)
    out = _loop_body(p, t, rect.eval, ef.eval, h0, 2)
    out.update(axis0=axes[0], axis1=axes[1], grid=grid, bbox=np.asarray(ef.bbox))
    np.savez_compressed(os.path.join(HERE, "loop_2d_grid.npz"), **out)

    # 3D gridded fh on the padded cube
    ef = size_fn_3d()
    cube = sm.Cube(ef.bbox)
    h0 = float(ef.hmin) * 2
    p, t = _jittered_mesh(cube, h0, ef.bbox, 9, 3)
    axes, grid = _rgi_parts(ef)
    out = _loop_body(p, t, cube.eval, ef.eval, h0, 3)
    out.update(axis0=axes[0], axis1=axes[1], axis2=axes[2], grid=grid, bbox=np.asarray(ef.bbox))
    np.savez_compressed(os.path.join(HERE, "loop_3d_grid.npz"), **out)


def gen_sliver():
    ball = sm.Ball([0.0, 0.0, 0.0], 1.0)
    h0 = 0.2
    p, t = _jittered_mesh(ball, h0, ball.bbox, 11, 3)
    t = mg._remove_triangles_outside(p, t, ball.eval, 0.1 * h0)
    dh = sm.geometry.calc_dihedral_angles(p, t)
    lo, hi = 10.0 * np.pi / 180, 180.0 * np.pi / 180
    ele, _ = mg._calc_dihedral_angles(p, t, lo, hi)
    s = t[ele]
    g = sm.geometry.calc_circumsphere_grad(p[s[:, 0]], p[s[:, 1]], p[s[:, 2]], p[s[:, 3]])
    perturb = g.copy()
    perturb[np.isinf(perturb)] = 1.0
    perturb /= (np.sum(np.abs(perturb) ** 2, axis=-1) ** 0.5)[:, None]
    pn = p.copy()
    step = 0.10
    pn[s[:, 0]] += step * h0 * perturb
    np.savez_compressed(
        os.path.join(HERE, "sliver_3d.npz"), p=p, t=t, dh=np.asarray(dh).ravel(), ele=ele, grad=g,
        p_new=pn, step=step, h0=h0, min_dh=lo, max_dh=hi,
    )


def gen_init():
    out = {}
    comm = sys.modules["mpi4py"].MPI.COMM_WORLD
    for name, dom, el, dim in (
        ("disk", sm.Disk([0.0, 0.0], 1.0), 0.05, 2),
        ("ball", sm.Ball([0.0, 0.0, 0.0], 1.0), 0.2, 3),
    ):
        fh, _, hmin, lsf = mg._unpack_sizing(el)
        bbox = np.array(dom.bbox).reshape(-1, 2)
        opts = dict(seed=0, r0m_is_h0=False, axis=1, points=None)
        _, p, _ = mg._generate_initial_points(el, 0.1 * el, dim, bbox, fh, dom.eval, np.empty((0, dim)), comm, opts, lsf)
        out[name] = p
    ef = size_fn_2d()
    rect = sm.Rectangle(ef.bbox)
    h0 = float(ef.hmin)
    bbox = np.array(ef.bbox).reshape(-1, 2)
    opts = dict(seed=0, r0m_is_h0=False, axis=1, points=None)
    pfix = rect.corners
    _, p, _ = mg._generate_initial_points(h0, 0.1 * h0, 2, bbox, ef.eval, rect.eval, pfix, comm, opts, True)
    out["grid2d"] = p
    np.savez_compressed(os.path.join(HERE, "init_points.npz"), **out)


def gen_e2e():
    res = {}
    q = sm.geometry.simp_qual
    p, t = sm.generate_mesh(sm.Disk([0.0, 0.0], 1.0), 0.05, max_iter=25, verbose=0)
    res["disk_h0.05"] = dict(nverts=len(p), ncells=len(t), min_q=float(q(p, t).min()), mean_q=float(q(p, t).mean()),
                             area=float(sm.geometry.simp_vol(p, t).sum()))
    ef = size_fn_2d()
    p, t = sm.generate_mesh(sm.Rectangle(ef.bbox), ef, max_iter=25, verbose=0)
    res["grid2d"] = dict(nverts=len(p), ncells=len(t), min_q=float(q(p, t).min()), mean_q=float(q(p, t).mean()),
                         area=float(sm.geometry.simp_vol(p, t).sum()))
    # quarter annulus (tests/test_2d_min_qual.py) at h=0.05
    rect = sm.Rectangle((0.0, 1.0, 0.0, 1.0))
    ann = sm.Intersection([rect, sm.Difference([sm.Disk([0.0, 0.0], 1.0), sm.Disk([0.0, 0.0], 0.5)])])
    p, t = sm.generate_mesh(ann, 0.05, max_iter=25, verbose=0)
    res["annulus_h0.05"] = dict(nverts=len(p), ncells=len(t), min_q=float(q(p, t).min()), mean_q=float(q(p, t).mean()),
                                area=float(sm.geometry.simp_vol(p, t).sum()))
    ball = sm.Ball([0.0, 0.0, 0.0], 1.0)
    p, t = sm.generate_mesh(ball, 0.2, max_iter=25, verbose=0)
    n_gen = len(p)
    p, t = sm.sliver_removal(points=p, domain=ball, edge_length=0.2, verbose=0)
    dh = sm.geometry.calc_dihedral_angles(p, t)
    res["ball_h0.2"] = dict(nverts_generate=n_gen, nverts=len(p), ncells=len(t), min_dihedral_deg=float(dh.min() * 180 / np.pi),
                            volume=float(sm.geometry.simp_vol(p, t).sum()))
    with open(os.path.join(HERE, "e2e.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


def gen_meshutil():
    """meshutil_2d.npz / meshutil_3d.npz: the reference's mesh utilities used at termination
    (geometry/utils.py: simp_vol :175, fix_mesh :204, simp_qual :252, get_boundary_edges :310,
    get_boundary_vertices :364, get_boundary_entities :385, get_boundary_facets :421,
    delete_boundary_entities :440, laplacian2_fixed_point :494) on small jittered meshes."""
    gu = sm.geometry.utils
    out = {}
    p, t = _jittered_mesh(sm.Disk([0.0, 0.0], 1.0), 0.12, (-1.0, 1.0, -1.0, 1.0), 5, 2)
    t = mg._remove_triangles_outside(p, t, sm.Disk([0.0, 0.0], 1.0).eval, 0.012)
    out["p"], out["t"] = p, t.astype(np.int64)
    out["vol"], out["qual"] = gu.simp_vol(p, t), gu.simp_qual(p, t)
    out["bedges"] = gu.get_boundary_edges(t)
    out["bverts"] = gu.get_boundary_vertices(t)
    out["bents"] = gu.get_boundary_entities(p, t)
    # a dirty mesh: duplicated vertices (cells re-pointed at the copies), a duplicated cell, an unused vertex
    pd = np.vstack((p, p[:7], [[5.0, 5.0]]))
    td = t.copy()
    td[::3][td[::3] < 7] += len(p)
    td = np.vstack((td, td[:4, [1, 2, 0]]))
    out["dirty_p"], out["dirty_t"] = pd, td.astype(np.int64)
    fp, ft, fj = gu.fix_mesh(pd.copy(), td.copy(), delete_unused=True)
    out["fix_p"], out["fix_t"] = fp, ft
    fp2, ft2, _ = gu.fix_mesh(pd.copy(), td.copy(), delete_unused=False)
    out["fix2_p"], out["fix2_t"] = fp2, ft2
    dp, dt_ = gu.delete_boundary_entities(p.copy(), t.copy(), dim=2, min_qual=0.55, verbose=0)
    out["del_p"], out["del_t"] = dp, dt_
    lp, _ = gu.laplacian2_fixed_point(p.copy(), t.copy())
    out["lap_p"] = lp
    np.savez(os.path.join(HERE, "meshutil_2d.npz"), **out)
    out = {}
    ball = sm.Ball([0.0, 0.0, 0.0], 1.0)
    p, t = _jittered_mesh(ball, 0.3, (-1.0, 1.0, -1.0, 1.0, -1.0, 1.0), 6, 3)
    t = mg._remove_triangles_outside(p, t, ball.eval, 0.03)
    out["p"], out["t"] = p, t.astype(np.int64)
    out["vol"], out["qual"] = gu.simp_vol(p, t), gu.simp_qual(p, t)
    out["bfacets"] = gu.get_boundary_facets(t)
    out["bverts"] = gu.get_boundary_vertices(t, dim=3)
    out["bents"] = gu.get_boundary_entities(p, t, dim=3)
    dp, dt_ = gu.delete_boundary_entities(p.copy(), t.copy(), dim=3, min_qual=0.2, verbose=0)
    out["del_p"], out["del_t"] = dp, dt_
    np.savez(os.path.join(HERE, "meshutil_3d.npz"), **out)
    print("meshutil goldens:", {k: np.asarray(v).shape for k, v in out.items()})


def gen_reference_tests():
    """reference_tests.json: the reference's own known-answer tests on this path, replayed with the
    unmodified reference (Qhull behind its CGAL interface), next to the answers those tests assert:
    tests/test_2dmesher_SDF.py, test_immersion.py, test_smooth_sets.py, test_pfix.py, test_verbose.py."""
    import contextlib
    import io

    res = {}
    with contextlib.redirect_stdout(io.StringIO()):
        disk = sm.geometry.Disk([0.0, 0.0], 1)
        p, c = sm.generate_mesh(bbox=(-1.0, 1.0, -1.0, 1.0), domain=disk, h0=0.2,
                                edge_length=lambda x: 0.2 - disk.eval(x) * 0.15, max_iter=100)
        res["test_2dmesher_SDF"] = {"asserted": {"counts": [63, 93], "atol": 10, "area": 3.14, "area_atol": 0.2},
                                    "reference_run_here": [len(p), len(c), float(sm.geometry.simp_vol(p, c).sum())]}
        runs = []
        for radius in [0.25, 0.30, 0.35]:
            disk0 = sm.Disk([0.5, 0.5], radius)
            p, c = sm.generate_mesh(domain=sm.Rectangle((0.0, 1.0, 0.0, 1.0)), h0=0.05, subdomains=[disk0],
                                    edge_length=lambda x, d=disk0: 0.05 * np.abs(d.eval(x)) + 0.05)
            sd = disk0.eval(p[c].sum(1) / 3)
            runs.append([radius, len(p), len(c), float(np.sum(sm.geometry.simp_vol(p, c[sd < 0])))])
        res["test_immersion"] = {"asserted": "immersed area == pi r^2 to rtol 1e-2", "reference_run_here": runs}
        dom = sm.Difference([sm.Ball((0.0, 0.0, 0.5), 0.85), sm.Cube((-0.5, 0.5, -0.5, 0.5, -0.5, 0.5))], smoothness=0.20)
        p, c = sm.generate_mesh(domain=dom, edge_length=0.10)
        p, c = sm.sliver_removal(points=p, domain=dom, edge_length=0.10)
        res["test_smooth_diff"] = {"asserted": {"cells": 9004, "atol": 100}, "reference_run_here": [len(p), len(c)]}
        bbox = (0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
        pf = np.vstack((np.linspace((0.0, 0.0, 0.0), (1.0, 0.0, 1.0), int(np.sqrt(2) / 0.05)), sm.geometry.corners(bbox)))
        p, c = sm.generate_mesh(domain=sm.Cube(bbox), edge_length=0.05, pfix=pf)
        res["test_pfix"] = {"asserted": "every fixed point is a mesh vertex (squared distance isclose 0)",
                            "reference_run_here": [len(p), len(c), max(float(((p - q) ** 2).sum(1).min()) for q in pf)]}
    sizes = []
    for v in (0, 1, 2):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            sm.generate_mesh(domain=sm.Rectangle((0.0, 1.0, 0.0, 1.0)), edge_length=0.1, verbose=v)
        sizes.append(len(buf.getvalue().encode()))
    res["test_verbose"] = {"asserted": {"stdout_bytes": [0, 192, 6014]}, "reference_run_here": sizes}
    with open(os.path.join(HERE, "reference_tests.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


def gen_segy():
    """segy_testing.npz / segy_tests.json: the velocity model of the reference's own SEG-Y fixture
    (tests/testing.segy, decoded value by value by the harness's segyio stand-in) and the outcome of
    the reference's tests that mesh it (tests/test_2dmesher_domain_extension.py:16-62 with the
    answers it asserts to +-100; tests/test_2dmesher.py:16-46), replayed with the unmodified reference."""
    fname = os.path.join(ref_harness.REF_ROOT, "tests", "testing.segy")
    with ref_harness._SegyFile(fname) as f:
        vp = np.zeros((len(f.samples), len(f.trace)))
        for k, tr in enumerate(f.trace):
            vp[:, k] = tr
    np.savez(os.path.join(HERE, "segy_testing.npz"), vp=np.flipud(vp), traces=vp)
    bbox = (-10e3, 0.0, 0.0, 10e3)
    res = {"bbox": bbox, "domain_extension": {}}
    for style, answer in (("linear_ramp", [9428, 18525]), ("edge", [9724, 19078]), ("constant", [9428, 18525])):
        ef = sm.get_sizing_function_from_segy(fname, bbox=bbox, grade=0.005, grad=50.0, stencil_size=100, wl=5, freq=5.0,
                                              hmin=100, hmax=10e6, pad_style=style, domain_pad=1e3)
        p, t = sm.generate_mesh(sm.Rectangle(bbox), ef, h0=100, perform_checks=True, verbose=0)
        res["domain_extension"][style] = {"asserted_by_reference_test": answer, "reference_run_here": [len(p), len(t)]}
    ef = sm.get_sizing_function_from_segy(fname, bbox=bbox, grade=0.005, grad=50.0, wl=5, freq=5.0, hmin=100, hmax=10e6)
    p, t = sm.generate_mesh(sm.Rectangle(bbox), ef, h0=100, perform_checks=True, verbose=0)
    res["test_2dmesher"] = {"reference_run_here": [len(p), len(t)]}
    with open(os.path.join(HERE, "segy_tests.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


def gen_reference_tests_3d():
    """bin3d_testing.npz / reference_tests_3d.json: the reference's 20x10x10 binary velocity fixture
    (tests/test3D.bin, float32 little-endian, kept as raw values) and the outcome of the reference tests that
    mesh it or exercise the 3-D / water-layer paths, replayed with the unmodified reference next to the answers
    those tests state: tests/test_3dmesher.py, test_3dmesher_domain_extension.py, test_3dmesher_SDF.py,
    test_2dmesher_vs_water.py."""
    import contextlib
    import io

    fname = os.path.join(ref_harness.REF_ROOT, "tests", "test3D.bin")
    np.savez(os.path.join(HERE, "bin3d_testing.npz"), raw=np.fromfile(fname, dtype="<f4"))
    res = {}
    bbox = (-2e3, 0.0, 0.0, 1e3, 0.0, 1e3)
    cube = sm.Cube(bbox)

    def dh_range(p, c):
        dh = np.asarray(sm.geometry.calc_dihedral_angles(p, c)).reshape(-1) * 180.0 / np.pi
        return [float(dh.min()), float(dh.max())]

    with contextlib.redirect_stdout(io.StringIO()):
        ef = sm.get_sizing_function_from_segy(fname, bbox, grade=0.005, grad=50, freq=2, wl=10, hmin=50, nz=20, nx=10, ny=10,
                                              byte_order="little", domain_pad=0.0, axes_order=(2, 0, 1))
        p, c = sm.generate_mesh(domain=cube, edge_length=ef, h0=50, max_iter=25, perform_checks=False)
        p, c = sm.sliver_removal(points=p, edge_length=ef, domain=cube, h0=50)
        res["test_3dmesher"] = {"stated_by_reference_test": [16459, 89240], "atol": 100, "reference_run_here": [len(p), len(c)],
                                "dihedral_range_deg": dh_range(p, c)}
        ext = {}
        for style, answer in (("linear_ramp", [1388, 6592]), ("edge", [1383, 6545]), ("constant", [1406, 6622])):
            ef = sm.get_sizing_function_from_segy(fname, bbox, grade=0.005, grad=150, freq=2, wl=5, hmin=150, nz=20, nx=10, ny=10,
                                                  byte_order="little", domain_pad=200, pad_style=style, axes_order=(2, 0, 1))
            p, c = sm.generate_mesh(domain=cube, edge_length=ef, h0=150, perform_checks=False)
            p, c = sm.sliver_removal(points=p, domain=cube, edge_length=ef, h0=150)
            ext[style] = {"stated_by_reference_test": answer, "reference_run_here": [len(p), len(c)], "dihedral_range_deg": dh_range(p, c)}
        res["test_3dmesher_domain_extension"] = {"atol": 100, "styles": ext}

        def cylinder(q):
            r, z = np.sqrt(q[:, 0] ** 2 + q[:, 1] ** 2), q[:, 2]
            d1, d2, d3 = r - 1.0, z - 1.0, -z - 1.0
            d4, d5 = np.sqrt(d1**2 + d2**2), np.sqrt(d1**2 + d3**2)
            d = np.maximum.reduce([d1, d2, d3])
            ix = (d1 > 0) * (d2 > 0)
            d[ix] = d4[ix]
            ix = (d1 > 0) * (d3 > 0)
            d[ix] = d5[ix]
            return d

        cb = (-1.0, 1.0, -1.0, 1.0, -1.0, 1.0)
        p, c = sm.generate_mesh(bbox=cb, domain=cylinder, h0=0.10, edge_length=lambda q: np.array([0.10] * len(q)), max_iter=100)
        p, c = sm.sliver_removal(points=p, domain=cylinder, edge_length=lambda q: np.array([0.10] * len(q)), h0=0.10, bbox=cb)
        res["test_3dmesher_SDF"] = {"asserted": {"volume": 6.28, "atol": 0.10}, "stated_counts": [6825, 36206],
                                    "reference_run_here": [len(p), len(c), float(np.sum(sm.geometry.simp_vol(p, c)))]}
        vs = np.zeros((200, 200))
        vs[0:150, :] = 1000
        rb = (-10000.0, 0.0, 0.0, 10000.0)
        ef = sm.get_sizing_function_from_segy(None, bbox=rb, grade=0.0, grad=0.0, wl=5, freq=2.0, hmin=10, hmax=10e6, velocity_data=vs,
                                              nz=200, nx=200)
        probes = [float(ef.eval((-5000, 5000))), float(ef.eval((-1, 5000)))]
        ef.hmin = None
        p, c = sm.generate_mesh(sm.Rectangle(rb), ef, h0=125, perform_checks=True)
        res["test_2dmesher_vs_water"] = {"asserted": {"probes": [100, 150], "counts": [5616, 10955], "atol": 100},
                                         "reference_run_here": {"probes": probes, "counts": [len(p), len(c)]}}
    with open(os.path.join(HERE, "reference_tests_3d.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if "reftests" in sys.argv[1:]:
        gen_reference_tests()
        sys.exit(0)
    if "reftests3d" in sys.argv[1:]:
        gen_reference_tests_3d()
        sys.exit(0)
    if "segy" in sys.argv[1:] or "meshutil" in sys.argv[1:]:
        if "segy" in sys.argv[1:]:
            gen_segy()
        if "meshutil" in sys.argv[1:]:
            gen_meshutil()
        sys.exit(0)
    gen_sdf()
    gen_interp()
    gen_loop()
    gen_sliver()
    gen_init()
    gen_e2e()
    gen_sizing()
    gen_segy()
    gen_meshutil()
    gen_reference_tests()
    gen_reference_tests_3d()
    print("golden vectors written to", HERE)
