"""``generate_mesh`` and ``sliver_removal`` with the reference's Python API
(SeismicMesh/generation/mesh_generator.py:59-529): same positional arguments, kwargs, defaults,
error types/messages and verbose output -- but the loop body (cull, unique bars, forces, update,
projection, convergence monitor; dihedral test and sliver perturbation) runs in CUDA kernels on
device-resident arrays.  Delaunay retriangulation stays on the host (``triangulator.py``) and its
time is reported separately in ``last_run_stats``.
"""
import math
import time
import warnings

import numpy as np
import torch

from . import device as D
from . import geometry, meshutil
from ._lib import check, lib
from .engine import ForceLoop, Level, SizeSpec, laplacian_smooth
from .sizing import SizeFunction
from .triangulator import get_triangulator

__all__ = ["generate_mesh", "sliver_removal", "last_run_stats"]

#: timing breakdown of the most recent call (seconds): delaunay (host), h2d, device loop, d2h,
#: opaque user callables, termination; plus vertex/iteration counts.
last_run_stats = {}

_ALLOWED = {
    "verbose", "max_iter", "seed", "perform_checks", "pfix", "axis", "points", "domain", "edge_length",
    "bbox", "min_dh_angle_bound", "max_dh_angle_bound", "delta_t", "h0", "geps_mult", "subdomains",
    "gamma", "preserve", "mesh_improvement", "r0m_is_h0",
}
# API extensions of this implementation (documented in DESIGN.md): none are needed for parity.
# `ttol`: opt-in DistMesh displacement test (Persson & Strang): retriangulate only when some vertex has
# moved more than ttol*h0 since the last Delaunay; iterations in between re-use the neighbour rows
# and never leave the device.  Default None = the reference's behaviour (retriangulate every iteration).
_EXTENSIONS = {"triangulator", "ttol"}


def _parse_kwargs(kwargs):
    for key in kwargs:
        if key not in _ALLOWED and key not in _EXTENSIONS:
            raise ValueError("Option %s with parameter %s not recognized " % (key, kwargs[key]))


def _printers(opts):
    v = opts["verbose"]
    if v < 0:
        raise ValueError("Unknown verbosity level")

    def p1(msg):
        if v >= 1:
            print(msg, flush=True)

    def p2(msg):
        if v > 1:
            print(msg, flush=True)

    return p1, p2


def _check_bbox(bbox):
    if bbox is not None:
        for b in bbox:
            if isinstance(b, int):
                raise ValueError("bbox must contain all floats")


def _minmax(b0, b1):
    return tuple(min(a, b) if i % 2 == 0 else max(a, b) for i, (a, b) in enumerate(zip(b0, b1)))


def _unpack_domain(domain, opts):
    """-> (domain object or callable, bbox, corners)   (reference :594-621)"""
    corners = None
    if isinstance(domain, geometry._SDF) or (hasattr(domain, "eval") and hasattr(domain, "bbox")):
        bbox = domain.bbox
        corners = getattr(domain, "corners", None)
    elif callable(domain):
        bbox = opts["bbox"]
    else:
        raise ValueError("`domain` must be a function or a :class:`geometry` object")
    _check_bbox(bbox)
    return domain, bbox, corners


def _unpack_sizing(edge_length, dim_hint=None):
    """-> (kind, payload, bbox, hmin)   (reference :562-591)"""
    if isinstance(edge_length, SizeFunction) or (
        hasattr(edge_length, "eval") and hasattr(edge_length, "bbox") and hasattr(edge_length, "hmin")
    ):
        bbox, hmin = edge_length.bbox, edge_length.hmin
        interp = edge_length.interpolant() if isinstance(edge_length, SizeFunction) else None
        payload = ("grid", interp) if interp is not None else ("func", edge_length.eval)
    elif callable(edge_length):
        bbox, hmin, payload = None, None, ("func", edge_length)
    elif np.isscalar(edge_length):
        bbox, hmin, payload = None, edge_length, ("const", float(edge_length))
    else:
        raise ValueError("`edge_length` must either be a function, a `edge_length` object, or a scalar")
    _check_bbox(bbox)
    return payload, bbox, hmin


def _size_spec(payload, dim):
    kind, val = payload
    if kind == "const":
        return SizeSpec(dim, const=val)
    if kind == "grid":
        return SizeSpec(dim, interp=val)
    return SizeSpec(dim, func=val)


def _staggered_grid(h0, dim, bbox):
    """Initial lattice of spacing h0 with every other row shifted by h0/2 (behaviour of
    generation/utils.py:15-25; bbox is (dim,2))."""
    # np.mgrid[slice(lo, hi + h0, h0)] yields lo + i*h0 for i < ceil((hi + h0 - lo) / h0)
    axes = [np.arange(int(math.ceil((hi + h0 - lo) / h0)), dtype=float) * h0 + lo for lo, hi in bbox]
    grids = [g.copy() for g in np.meshgrid(*axes, indexing="ij")]
    # odd planes along the FIRST axis are shifted in the second (and, in 3-D, also the third)
    # coordinate -- the reference indexes axis 0 for both shifts (generation/utils.py:19-23)
    grids[1][1::2] += h0 / 2
    if dim == 3:
        grids[2][1::2] += h0 / 2
    return np.stack([g.ravel() for g in grids], axis=1)


def _initial_points(h0, geps, dim, bbox, size, level0, pfix, opts):
    """Lattice -> keep fd<geps -> probabilistic rejection r0m^d/r0^d with NumPy's legacy RNG
    (reference :808-852; the RNG stream decides the vertex count, so it stays NumPy on the host)."""
    p = _staggered_grid(h0, dim, bbox)
    p = p[level0.eval_host(p) < geps]
    r0 = size.eval_host(p)
    if opts["r0m_is_h0"]:
        alpha = 1.1
        if alpha * h0 < r0.min():
            r0m = alpha * h0
            if dim == 3:
                warnings.warn("Warning: r0m_is_h0 option not test for 3D meshes.")
        else:
            r0m = h0
    else:
        r0m = r0.min()
    np.random.seed(opts["seed"])
    p = p[np.random.rand(p.shape[0]) < r0m**dim / r0**dim]
    if len(pfix) and len(p):
        # A lattice point that coincides with a fixed point (the lattice origin and box corners ARE the
        # corners of a Rectangle / Cube) would be an exact duplicate row: a Delaunay code keeps ONE copy,
        # not necessarily the fixed row, and the constraint would silently move to a free vertex.  The
        # reference gets away with it because CGAL merges duplicates and pfix is re-located by
        # nearest-node search every iteration (mesh_generator.py:472-475); here the lattice copy goes.
        # (The random stream above is drawn for the full lattice first, so the vertex set is the
        # reference's minus those copies.)
        _, d2 = _nearest_sq(pfix, p)
        p = p[d2 > 0.0]
    return np.vstack((pfix, p))


def _nearest_sq(a, b):
    """for every row of b: (index, squared distance) of the nearest row of a (a is small)."""
    best = np.full(len(b), np.inf)
    arg = np.zeros(len(b), dtype=np.int64)
    for i, row in enumerate(np.asarray(a, dtype=np.float64)):
        d = ((b - row) ** 2).sum(axis=1)
        m = d < best
        best[m] = d[m]
        arg[m] = i
    return arg, best


def _termination(p, t, opts, dim, sliver=False, verbose=1):
    """Host clean-up at max_iter (reference :655-677)."""
    if dim == 2:
        p, t, _ = meshutil.fix_mesh(p, t, dim=dim, delete_unused=True)
        p, t = meshutil.delete_boundary_entities(p, t, dim=2, min_qual=0.15, verbose=verbose)
        if opts["subdomains"] is None and opts["mesh_improvement"]:
            p, t = laplacian_smooth(p, t)  # the linear solve of geometry.laplacian2_fixed_point, on the device
    if opts["perform_checks"]:  # reference :672-675
        p, t = meshutil.linter(p, t, dim=dim)
    else:
        p, t, _ = meshutil.fix_mesh(p, t, dim=dim, delete_unused=True)
    return p, t


def _level_set_newton(p_host, t_host, level, deps, dim):
    """_improve_level_set_newton (reference :741-759): boundary vertices from the host, five
    damped Newton steps in one kernel."""
    bid = meshutil.get_boundary_vertices(t_host, dim)
    if len(bid) == 0:
        return p_host
    if level.lowered:
        pd = D.to_dev(p_host, torch.float64)
        bd = D.to_dev(bid.astype(np.int32), torch.int32)
        check(lib.dm_level_set_newton(D.ptr(level.prog), D.ptr(pd), D.ptr(bd), len(bid), dim, deps, D.stream_ptr()), "level_set_newton")
        return pd.cpu().numpy()
    p = p_host.copy()
    fd = level.func
    alpha = 1
    for iteration in range(5):
        d = fd(p[bid])
        grads = []
        for i in range(dim):
            dv = np.zeros(dim)
            dv[i] = deps
            grads.append((fd(p[bid] + dv) - d) / deps)
        g2 = sum(g**2 for g in grads)
        g2 = np.where(g2 < deps, deps, g2)
        p[bid] -= alpha * (d * np.vstack(grads) / g2).T
        alpha /= iteration + 1
    return p


def _comm_info(comm):
    if comm is None:
        return 0, 1
    return int(getattr(comm, "rank", 0)), int(getattr(comm, "size", 1))


def _finish(loop, p_host, p_dev, t_dev, level0, size, h0, deps, dim, gen_opts, stats, print_msg1):
    """max_iter reached (reference :485-494): cull, host clean-up, level-set Newton, final tight cull."""
    print_msg1("Termination reached...maximum number of iterations reached.")
    tt = time.perf_counter()
    t_kept = loop.kept_cells(p_dev, t_dev).cpu().numpy()
    p_out, t_out = _termination(p_host, t_kept, gen_opts, dim, verbose=gen_opts["verbose"])
    p_out = _level_set_newton(p_out, t_out, level0, deps, dim)
    # final cull with a tight tolerance (reference :493)
    pd = D.to_dev(p_out, torch.float64)
    td = D.to_dev(t_out, torch.int32)
    fin = ForceLoop(dim, [level0], size, h0, h0 * 0.001, deps)
    t_out = fin.kept_cells(pd, td).cpu().numpy().astype(t_out.dtype)
    stats["termination"] += time.perf_counter() - tt
    return p_out, t_out


def _loop_pinned(loop, tri, p, N, dim, h0, deps, max_iter, ttol, level0, size, gen_opts, stats, print_msg1, print_msg2):
    """The DistMesh loop (reference :460-527) for lowered fd / fh: positions live on the device, the host
    Delaunay reads them from / writes its cells to PINNED staging buffers (raw-pointer triangulators,
    no intermediate arrays), the cell list goes up in chunks overlapped with stage A and the new
    positions come back asynchronously (ForceLoop.iterate_host -- the call bench.py's `e2e` leg times).
    `ttol` (extension): iterations between two retriangulations re-use the neighbour rows and never
    leave the device; the displacement test is relative to the local mesh size."""
    pins = [torch.empty((N, dim), dtype=torch.float64).pin_memory() for _ in range(2)]
    pins[0].copy_(torch.from_numpy(np.ascontiguousarray(p)))
    cur = 0                       # pins[cur] mirrors the device positions whenever host_current
    cap = max(int(tri.max_cells(N)) if hasattr(tri, "max_cells") else 8 * N + 64, 1)
    t_pin = torch.empty((cap, dim + 1), dtype=torch.int32).pin_memory()
    sc_pin = torch.empty(8, dtype=torch.float64).pin_memory()
    p_dev = D.to_dev(p, torch.float64)
    host_current = True
    p_tri = None                  # positions at the last retriangulation (ttol test)
    t_dev = None
    T = 0
    disp = float("inf")
    count = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    while True:
        start = time.time()
        last = count == (max_iter - 1)
        retri = last or ttol is None or p_tri is None or disp > ttol
        if retri:
            if not host_current:
                t3 = time.perf_counter()
                pins[cur].copy_(p_dev, non_blocking=True)
                torch.cuda.synchronize()
                host_current = True
                stats["d2h"] += time.perf_counter() - t3
            t0 = time.perf_counter()
            pts = pins[cur].numpy()
            if hasattr(tri, "triangulate_into"):
                T = tri.triangulate_into(pts, t_pin.numpy())
                if T < 0:  # capacity: a pathological point set needs more than the usual bound
                    t_pin = torch.empty((int(-T * 1.1) + 64, dim + 1), dtype=torch.int32).pin_memory()
                    T = tri.triangulate_into(pts, t_pin.numpy())
            else:
                th = np.ascontiguousarray(tri.triangulate(pts), dtype=np.int32)
                T = len(th)
                if T > t_pin.shape[0]:
                    t_pin = torch.empty((int(T * 1.1) + 64, dim + 1), dtype=torch.int32).pin_memory()
                t_pin.numpy()[:T] = th
            stats["delaunay"] += time.perf_counter() - t0
            stats["triangulations"] += 1
            if ttol is not None:
                p_tri = p_dev.clone()

        if last:
            t_dev = D.to_dev(t_pin[:T], torch.int32)
            return _finish(loop, pins[cur].numpy().copy(), p_dev, t_dev, level0, size, h0, deps, dim, gen_opts, stats,
                           print_msg1)

        ev0.record()
        if retri:
            nxt = cur ^ 1
            p_dev = loop.iterate_host(None, t_pin[:T], pins[nxt], p_dev=p_dev)
            cur = nxt
            host_current = True
        else:
            p_dev, _ = loop.iterate_reuse(p_dev, p_out=loop.spare_like(p_dev))
            host_current = False
        sc_pin.copy_(loop.plan.scalars(), non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        stats["device"] += ev0.elapsed_time(ev1) * 1e-3
        maxdp = float(sc_pin[4])
        if ttol is not None:
            disp = loop.displacement(p_dev, p_tri, relative=True)
        stats["iterations"] += 1
        print_msg2(
            "Iteration #%d, max movement is %f, there are %d vertices and %d cells"
            % (count + 1, maxdp, N, _kept_count(loop) if gen_opts["verbose"] > 1 else 0)
        )
        assert maxdp < 1000 * h0, "max movement indicates there's a convergence problem"
        count += 1
        print_msg2("     Elapsed wall-clock time %f : " % (time.time() - start))


def _loop_staged(loop, tri, p, N, dim, h0, deps, max_iter, level0, size, gen_opts, stats, print_msg1, print_msg2):
    """The same loop when `domain` / `edge_length` is an opaque Python callable: user code is evaluated
    on the host through staged copies inside ForceLoop.iterate (reported as `host_callables`)."""
    p_dev = D.to_dev(p, torch.float64)
    p_host = p
    count = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    while True:
        start = time.time()
        t0 = time.perf_counter()
        t_host = tri.triangulate(p_host)
        t1 = time.perf_counter()
        t_dev = D.to_dev(t_host, torch.int32)
        torch.cuda.synchronize()
        stats["delaunay"] += t1 - t0
        stats["h2d"] += time.perf_counter() - t1
        stats["triangulations"] += 1
        if count == (max_iter - 1):
            return _finish(loop, p_host, p_dev, t_dev, level0, size, h0, deps, dim, gen_opts, stats, print_msg1)
        ev0.record()
        p_dev, _ = loop.iterate(p_dev, t_dev)
        ev1.record()
        torch.cuda.synchronize()
        stats["device"] += ev0.elapsed_time(ev1) * 1e-3
        t3 = time.perf_counter()
        p_host = p_dev.cpu().numpy()
        maxdp = loop.maxdp()
        stats["d2h"] += time.perf_counter() - t3
        stats["iterations"] += 1
        print_msg2(
            "Iteration #%d, max movement is %f, there are %d vertices and %d cells"
            % (count + 1, maxdp, N, _kept_count(loop) if gen_opts["verbose"] > 1 else 0)
        )
        assert maxdp < 1000 * h0, "max movement indicates there's a convergence problem"
        count += 1
        print_msg2("     Elapsed wall-clock time %f : " % (time.time() - start))


def generate_mesh(domain, edge_length, comm=None, **kwargs):  # noqa: C901
    r"""Generate a 2D/3D simplicial mesh with DistMesh (see the reference docstring,
    mesh_generator.py:291-342, for the meaning of every keyword argument).

    :return: points (N,dim) float64, cells (T,dim+1) integer -- NumPy arrays.
    """
    rank, size_ = _comm_info(comm)
    if size_ > 1:
        from .parallel import generate_mesh_parallel

        return generate_mesh_parallel(domain, edge_length, comm, **kwargs)
    gen_opts = {
        "verbose": 1, "max_iter": 50, "seed": 0, "perform_checks": False, "pfix": None, "axis": 1,
        "points": None, "delta_t": 0.30, "geps_mult": 0.1, "subdomains": None, "mesh_improvement": True,
        "r0m_is_h0": False, "triangulator": None, "ttol": None,
    }
    gen_opts.update(kwargs)
    _parse_kwargs(kwargs)
    print_msg1, print_msg2 = _printers(gen_opts)

    dom, bbox0, corners = _unpack_domain(domain, gen_opts)
    payload, bbox1, hmin = _unpack_sizing(edge_length)
    bbox = bbox0 if bbox1 is None else _minmax(bbox0, bbox1)
    if not isinstance(bbox, tuple):
        raise ValueError("`bbox` must be a tuple")
    dim = int(len(bbox) / 2)
    if not np.isscalar(edge_length) and dim == 3:
        corners = None
    if bbox0 != bbox1 and bbox1 is not None:  # padded sizing domain: mesh the padded box
        dom = geometry.Rectangle(bbox) if dim == 2 else geometry.Cube(bbox)
    bbox_arr = np.array(bbox).reshape(-1, 2)

    h0 = hmin if hmin is not None else gen_opts["h0"]
    if h0 < 0:
        raise ValueError("`h0` must be > 0")
    delta_t = gen_opts["delta_t"]
    geps = gen_opts["geps_mult"] * h0
    deps = np.sqrt(np.finfo(np.double).eps) * h0

    level0 = Level(dom, dim)
    size = _size_spec(payload, dim)

    if gen_opts["pfix"] is not None:
        pfix = np.array(gen_opts["pfix"], dtype="d")
    else:
        pfix = np.empty((0, dim))
    if corners is not None:
        corners = corners[level0.eval_host(corners) > -1000 * deps]
        pfix = np.append(pfix, corners, axis=0)
    nfix = len(pfix)
    print_msg1(f"Constraining {nfix} fixed points...")

    if gen_opts["points"] is None:
        p = _initial_points(h0, geps, dim, bbox_arr, size, level0, pfix, gen_opts)
        nfix_dev = nfix
    else:
        p = np.ascontiguousarray(gen_opts["points"], dtype=np.float64)
        nfix_dev = 0

    if gen_opts["max_iter"] < 0:
        raise ValueError("`max_iter` must be > 0")
    max_iter = gen_opts["max_iter"]
    N = p.shape[0]
    assert N > 0, "No vertices to mesh with!"
    print_msg1("Commencing mesh generation with %d vertices on rank %d." % (N, rank))

    levels = [level0]
    if gen_opts["subdomains"] is not None:
        for sub in gen_opts["subdomains"]:
            sd, _, _ = _unpack_domain(sub, gen_opts)
            levels.append(Level(sd, dim))

    tri = get_triangulator(gen_opts["triangulator"], dim)
    loop = ForceLoop(dim, levels, size, h0, geps, deps, delta_t=delta_t, nfix=0)
    fixed_mask = None
    if gen_opts["points"] is not None and nfix > 0:
        # user-supplied points: locate the fixed points by nearest node (reference :472-475)
        ifix = [int(np.argmin(((p - f) ** 2).sum(1))) for f in pfix]
        fixed_mask = np.zeros(N, dtype=np.uint8)
        fixed_mask[ifix] = 1
    loop.nfix = nfix_dev
    loop.fixed_mask = None if fixed_mask is None else D.to_dev(fixed_mask, torch.uint8)

    ttol = gen_opts["ttol"]
    if ttol is not None:
        if ttol < 0:
            raise ValueError("`ttol` must be >= 0")
        if not loop.all_lowered:
            warnings.warn("`ttol` needs lowered `domain` / `edge_length` (no opaque callables); retriangulating every iteration")
            ttol = None
    stats = dict(delaunay=0.0, h2d=0.0, device=0.0, d2h=0.0, termination=0.0, iterations=0, nverts=N,
                 triangulator=tri.name, triangulations=0)
    if loop.all_lowered:
        p_host, t_host = _loop_pinned(loop, tri, p, N, dim, h0, deps, max_iter, ttol, level0, size, gen_opts, stats,
                                      print_msg1, print_msg2)
    else:
        p_host, t_host = _loop_staged(loop, tri, p, N, dim, h0, deps, max_iter, level0, size, gen_opts, stats,
                                      print_msg1, print_msg2)
    stats["host_callables"] = loop.host_seconds
    last_run_stats.clear()
    last_run_stats.update(stats)
    return p_host, t_host


def _kept_count(loop):
    T = int(loop.plan.c.T)
    return int(loop.plan.keep()[:T].sum().item())


def sliver_removal(points, domain, edge_length, comm=None, **kwargs):  # noqa: C901
    r"""Improve an existing 3D mesh by removing degenerate cells ("slivers"); keyword arguments
    as in the reference (mesh_generator.py:59-100).  Serial only, like the reference."""
    rank, _ = _comm_info(comm)
    if rank > 0:
        return True, True
    sliver_opts = {
        "verbose": 1, "max_iter": 50, "perform_checks": False, "axis": 1, "min_dh_angle_bound": 10.0,
        "max_dh_angle_bound": 180.0, "points": None, "delta_t": 0.30, "geps_mult": 0.1, "subdomains": [],
        "gamma": 1.0, "preserve": False, "triangulator": None,
    }
    sliver_opts.update(kwargs)
    _parse_kwargs(kwargs)
    print_msg1, print_msg2 = _printers(sliver_opts)

    dim = points.shape[1]
    if dim == 2:
        raise Exception("Mesh improvement currently on works in 3D")

    dom, bbox0, _ = _unpack_domain(domain, sliver_opts)
    payload, bbox1, hmin = _unpack_sizing(edge_length)
    bbox = bbox0 if bbox1 is None else _minmax(bbox0, bbox1)
    if bbox0 != bbox1 and bbox1 is not None:
        dom = geometry.Cube(bbox)
    if not isinstance(bbox, tuple):
        raise ValueError("`bbox` must be a tuple")
    h0 = hmin if hmin is not None else sliver_opts["h0"]
    if h0 < 0:
        raise ValueError("`h0` must be > 0")
    if sliver_opts["max_iter"] < 0:
        raise ValueError("`max_iter` must be > 0")
    max_iter = sliver_opts["max_iter"]
    print_msg1(f"Will attempt {max_iter} iterations to bound the dihedral angles...")
    geps = sliver_opts["geps_mult"] * h0
    deps = np.sqrt(np.finfo(np.double).eps) * h0
    min_dh_bound = sliver_opts["min_dh_angle_bound"] * math.pi / 180
    max_dh_bound = sliver_opts["max_dh_angle_bound"] * math.pi / 180
    print_msg1(f"Enforcing a min. dihedral bound of: {min_dh_bound * 180 / math.pi} degrees...")
    print_msg1(f"Enforcing a max. dihedral bound of: {max_dh_bound * 180 / math.pi} degrees...")

    level0 = Level(dom, dim)
    size = _size_spec(payload, dim)
    tri = get_triangulator(sliver_opts["triangulator"], dim)
    loop = ForceLoop(dim, [level0], size, h0, geps, deps)

    p_host = np.ascontiguousarray(points, dtype=np.float64)
    N = len(p_host)
    print_msg1("Commencing sliver removal with %d vertices on rank %d." % (N, rank))
    p_dev = D.to_dev(p_host, torch.float64)
    winner = torch.empty(N, dtype=torch.int32, device=p_dev.device)
    stats = dict(delaunay=0.0, device=0.0, iterations=0, nverts=N, triangulator=tri.name)
    count = 0
    step = 0.10
    gamma = sliver_opts["gamma"]
    num_old_bad = np.inf
    st = D.stream_ptr
    while True:
        start = time.time()
        t0 = time.perf_counter()
        t_host = tri.triangulate(p_host)
        stats["delaunay"] += time.perf_counter() - t0
        t_dev = D.to_dev(t_host, torch.int32)
        # which vertex sits in column 0 decides which vertex a sliver moves (reference :234,245-274 takes
        # CGAL's order): prefer a vertex well inside the domain, see dm_cells_lead_interior
        if level0.lowered:
            key = torch.empty(N, dtype=torch.float64, device=p_dev.device)
            check(lib.dm_sdf_eval(D.ptr(level0.prog), D.ptr(p_dev), N, dim, D.ptr(key), st()), "sdf_eval")
        else:
            key = D.to_dev(np.asarray(level0.func(p_host), dtype=np.float64), torch.float64)
        check(lib.dm_cells_lead_interior(D.ptr(key), D.ptr(t_dev), t_dev.shape[0], -0.5 * h0, st()), "cells_lead_interior")
        if level0.lowered:
            # cull + dihedral bound test in ONE kernel on the uncompacted cell list; the kept cells are
            # only compacted when the loop ends (fix_mesh below)
            t_kept = None
            t_use = t_dev
            T = t_dev.shape[0]
            flags = torch.empty(T, dtype=torch.uint8, device=p_dev.device)
            check(lib.dm_sliver_flags(D.ptr(level0.prog), D.ptr(p_dev), D.ptr(t_dev), T, geps, min_dh_bound, max_dh_bound,
                                      None, D.ptr(flags), st()), "sliver_flags")
            ele_nums = torch.nonzero(flags).flatten().to(torch.int32).cpu().numpy()
        else:
            t_kept = loop.kept_cells(p_dev, t_dev)
            t_use = t_kept
            T = t_kept.shape[0]
            flags = torch.empty(T, dtype=torch.uint8, device=p_dev.device)
            check(lib.dm_dihedral(D.ptr(p_dev), D.ptr(t_kept), T, min_dh_bound, max_dh_bound, None, D.ptr(flags), st()), "dihedral")
            ele_nums = np.nonzero(flags.cpu().numpy())[0].astype(np.int32)

        if count == (max_iter - 1):
            print_msg1(
                "FAILURE: Termination...maximum number of iterations reached. Try increasing max_iter when generating the mesh",
            )
            t_fin = t_kept if t_kept is not None else loop.kept_cells(p_dev, t_dev)
            p_host, t_out, _ = meshutil.fix_mesh(p_host, t_fin.cpu().numpy(), dim=dim, delete_unused=True)
            break
        print_msg1(f"On rank: {rank}. There are {len(ele_nums)} slivers...")
        if len(ele_nums) == 0:
            print_msg1(f"Termination reached in {count} iterations...no slivers detected!")
            t_fin = t_kept if t_kept is not None else loop.kept_cells(p_dev, t_dev)
            p_host, t_out, _ = meshutil.fix_mesh(p_host, t_fin.cpu().numpy(), dim=dim, delete_unused=True)
            break

        num_bad = len(ele_nums)
        if num_bad < num_old_bad:
            step /= gamma
        elif num_bad > num_old_bad:
            step *= gamma
        else:
            step /= 0.8  # stuck: increase the step
        ele_dev = D.to_dev(ele_nums, torch.int32)
        delta = torch.empty((num_bad, 3), dtype=torch.float64, device=p_dev.device)
        check(
            lib.dm_sliver_perturb(D.ptr(p_dev), N, D.ptr(t_use), D.ptr(ele_dev), num_bad, step * h0, D.ptr(winner), D.ptr(delta), st()),
            "sliver_perturb",
        )
        if sliver_opts["preserve"]:
            t_fin = t_kept if t_kept is not None else loop.kept_cells(p_dev, t_dev)
            ph = _level_set_newton(p_dev.cpu().numpy(), t_fin.cpu().numpy(), level0, deps, dim)
            p_dev = D.to_dev(ph, torch.float64)
        p_host = p_dev.cpu().numpy()
        count += 1
        stats["iterations"] += 1
        num_old_bad = num_bad
        print_msg2("     Elapsed wall-clock time %f : " % (time.time() - start))

    last_run_stats.clear()
    last_run_stats.update(stats)
    return p_host, t_out
