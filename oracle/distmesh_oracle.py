"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of the reference's DistMesh hot path.

This module is the *oracle* the CUDA path is checked against.  It is imported only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs; the product package (``seismicmesh_b200``) never imports it.

Every function cites the reference lines (relative to /root/reference/) it restates.  The
restatement is pinned against the reference's own code by ``tests/golden/make_golden.py``
(which imports the unmodified reference through ``oracle/ref_harness.py``) and by
``tests/test_oracle_vs_golden.py``; parity status: PINNED.

Third-party arithmetic on the path that is not under /root/reference:
  * scipy ``RegularGridInterpolator`` (scipy unpinned in the reference's setup.cfg:28-34,
    1.18.1 installed here) -> restated in :func:`interp_grid`, checked bit-for-bit against
    scipy itself in the tests.
  * scipy ``coo_matrix.toarray`` sequential accumulation -> restated in
    :func:`scatter_forces`.
"""
import math

import numpy as np

EPS = np.finfo(np.float64).eps

# ----------------------------------------------------------------------------------------
# SDF tree: spec = nested tuples
#   ("disk",  dict(x0=[xc,yc], r=r, rotate=[a,0,0], stretch=None|[..], translate=None|[..]))
#   ("ball",  dict(x0=[..3], r=r, ...)), ("rectangle", dict(bbox=(x1,x2,y1,y2), ...)),
#   ("cube", dict(bbox=(..6), ...)), ("torus", dict(r1=, r2=, ...)), ("prism", dict(b=, h=, ...)),
#   ("cylinder", dict(h=, r=, ...))
#   ("union"|"intersection"|"difference", [children...], k)
#   ("repeat", bbox6, child, period3)
# ----------------------------------------------------------------------------------------


def _manipulate(prm, x, dim):
    """geometry/signed_distance_functions.py:89-128 (translate, rotate back, scale back)."""
    t = prm.get("translate")
    if t is not None:
        x = x - np.asarray(t, dtype=np.float64)  # :123-124
    rot = prm.get("rotate") or [0.0, 0.0, 0.0]
    if dim == 2:
        a = rot[0]
        if a != 0.0:  # :93-94, R^T with R=[[c,-s],[s,c]] (:167-173)
            c, s = np.cos(a), np.sin(a)
            x = np.column_stack((c * x[:, 0] + s * x[:, 1], -s * x[:, 0] + c * x[:, 1]))
    else:
        if rot[0] != 0.0:  # Rx^T (:185-191)
            c, s = np.cos(rot[0]), np.sin(rot[0])
            x = np.column_stack((x[:, 0], c * x[:, 1] + s * x[:, 2], -s * x[:, 1] + c * x[:, 2]))
        if rot[1] != 0.0:  # Ry^T (:192-198)
            c, s = np.cos(rot[1]), np.sin(rot[1])
            x = np.column_stack((c * x[:, 0] - s * x[:, 2], x[:, 1], s * x[:, 0] + c * x[:, 2]))
        if rot[2] != 0.0:  # Rz^T (:199-205)
            c, s = np.cos(rot[2]), np.sin(rot[2])
            x = np.column_stack((c * x[:, 0] + s * x[:, 1], -s * x[:, 0] + c * x[:, 1], x[:, 2]))
    v = prm.get("stretch")
    if v is not None:  # :111-120 with v normalised in _build_stretch (:136-138)
        v = np.asarray(v, dtype=np.float64)
        alpha = np.sqrt(np.dot(v, v))
        vh = v / alpha
        dot = x[:, 0] * vh[0]
        for k in range(1, dim):
            dot = dot + x[:, k] * vh[k]
        vx = dot[:, None] * vh[None, :]
        x = vx / alpha + (x - vx)
    return x


def _length(x):
    """signed_distance_functions.py:70-71"""
    return np.sum(np.abs(x) ** 2, axis=-1) ** (1.0 / 2)


def sdf(spec, x):
    """Evaluate an SDF tree at x (M, dim).  signed_distance_functions.py:283-623 and the
    natives drectangle_fast / dblock_fast (geometry/cpp/fast_geometry.cpp:104-214)."""
    x = np.asarray(x, dtype=np.float64)
    kind = spec[0]
    if kind == "disk":  # :429-444, :596-598
        prm = spec[1]
        q = _manipulate(prm, x, 2)
        xc, yc = prm["x0"]
        return np.sqrt(((q - np.array([xc, yc])) ** 2).sum(-1)) - prm["r"]
    if kind == "ball":  # :450-468, :601-603
        prm = spec[1]
        q = _manipulate(prm, x, 3)
        xc, yc, zc = prm["x0"]
        return np.sqrt((q[:, 0] - xc) ** 2 + (q[:, 1] - yc) ** 2 + (q[:, 2] - zc) ** 2) - prm["r"]
    if kind == "rectangle":  # :474-489 -> fast_geometry.cpp:165-185
        prm = spec[1]
        q = _manipulate(prm, x, 2)
        x1, x2, y1, y2 = prm["bbox"]
        m = np.minimum
        return -m(m(m(-y1 + q[:, 1], y2 - q[:, 1]), -x1 + q[:, 0]), x2 - q[:, 0])
    if kind == "cube":  # :495-516 -> fast_geometry.cpp:104-134
        prm = spec[1]
        q = _manipulate(prm, x, 3)
        x1, x2, y1, y2, z1, z2 = prm["bbox"]
        m = np.minimum
        return -m(
            m(m(m(m(-z1 + q[:, 2], z2 - q[:, 2]), -y1 + q[:, 1]), y2 - q[:, 1]), -x1 + q[:, 0]),
            x2 - q[:, 0],
        )
    if kind == "torus":  # :522-540
        prm = spec[1]
        q = _manipulate(prm, x, 3)
        xz = np.column_stack((q[:, 0], q[:, 2]))
        qq = np.column_stack((_length(xz) - prm["r1"], q[:, 1]))
        return _length(qq) - prm["r2"]
    if kind == "prism":  # :546-563 (literal 0.866025, signed x1)
        prm = spec[1]
        q = _manipulate(prm, x, 3)
        a = np.abs(q)
        return np.maximum(
            a[:, 2] - prm["h"],
            np.maximum(a[:, 0] * 0.866025 + q[:, 1] * 0.5, -q[:, 1]) - prm["b"] * 0.5,
        )
    if kind == "cylinder":  # :569-590 (constructor halves h at :572)
        prm = spec[1]
        q = _manipulate(prm, x, 3)
        hh = prm["h"] / 2.0
        xz = np.column_stack((q[:, 0], q[:, 2]))
        lxz = np.column_stack((_length(xz), q[:, 1]))
        d = np.abs(lxz) - (prm["r"], hh)
        return np.minimum(np.maximum(d[:, 0], d[:, 1]), 0.0) + _length(np.maximum(d, 0.0))
    if kind in ("union", "intersection", "difference"):
        children, k = spec[1], spec[2]
        d = [sdf(c, x) for c in children]
        if kind == "union":  # :332-341
            if k == 0.0:
                return np.minimum.reduce(d)
            acc = d[0]
            for b in d[1:]:
                h = np.maximum(k - np.abs(acc - b), 0.0)
                acc = np.minimum(acc, b) - np.divide(h * h * 0.25, k)
            return acc
        if kind == "intersection":  # :372-381
            if k == 0.0:
                return np.maximum.reduce(d)
            acc = d[0]
            for b in d[1:]:
                h = np.maximum(k - np.abs(acc - b), 0.0)
                acc = np.maximum(acc, b) + h * h * 0.25 / k
            return acc
        # difference :412-423 (smooth variant folds the REVERSED child list)
        if k == 0.0:
            return np.maximum.reduce([-v if n > 0 else v for n, v in enumerate(d)])
        d = d[::-1]
        acc = d[0]
        for b in d[1:]:
            h = np.maximum(k - np.abs(-acc - b), 0.0)
            acc = np.maximum(-acc, b) + np.divide(h * h * 0.25, k)
        return acc
    if kind == "repeat":  # :283-294 (np.mod = floored modulo)
        bbox, child, period = spec[1], spec[2], np.asarray(spec[3], dtype=np.float64)
        q = np.mod(x + 0.5 * period, period) - 0.5 * period
        parent = ("cube", dict(bbox=tuple(bbox)))
        return np.maximum(sdf(child, q), sdf(parent, x))
    raise ValueError(f"unknown sdf node {kind}")


# ----------------------------------------------------------------------------------------
# gridded mesh-size function
# ----------------------------------------------------------------------------------------


def grid_axes(bbox, shape):
    """sizing/mesh_size_function.py:514-523: float32 linspace axes (scipy upcasts to f64)."""
    return [
        np.linspace(bbox[2 * k], bbox[2 * k + 1], n, dtype=np.float32).astype(np.float64)
        for k, n in enumerate(shape)
    ]


def interp_grid(axes, grid, x):
    """scipy RegularGridInterpolator(method='linear', bounds_error=False, fill_value=None)
    as built at sizing/mesh_size_function.py:391-408 and called via size_function.py:11-12.

    Restated from scipy 1.18 `_rgi.py` (find_indices + evaluate_linear_2d Cython fast path for
    2-D float64 grids; `_evaluate_linear` product-order loop for 3-D).  SURVEY.md section 3.5.
    """
    x = np.asarray(x, dtype=np.float64)
    dim = len(axes)
    idx, nrm = [], []
    for k in range(dim):
        a = axes[k]
        i = np.clip(np.searchsorted(a, x[:, k], side="right") - 1, 0, len(a) - 2)
        idx.append(i)
        nrm.append((x[:, k] - a[i]) / (a[i + 1] - a[i]))
    if dim == 2:
        i0, i1 = idx
        y0, y1 = nrm
        out = grid[i0, i1] * (1 - y0) * (1 - y1)
        out = out + grid[i0, i1 + 1] * (1 - y0) * y1
        out = out + grid[i0 + 1, i1] * y0 * (1 - y1)
        out = out + grid[i0 + 1, i1 + 1] * y0 * y1
        return out
    out = np.zeros(len(x))
    for c0 in (0, 1):
        for c1 in (0, 1):
            for c2 in (0, 1):
                w = 1.0
                for c, y in zip((c0, c1, c2), nrm):
                    w = w * (y if c else (1 - y))
                out = out + grid[idx[0] + c0, idx[1] + c1, idx[2] + c2] * w
    return out


# ----------------------------------------------------------------------------------------
# hot loop stages
# ----------------------------------------------------------------------------------------


def centroids(p, t):
    """generation/mesh_generator.py:737 : p[t].sum(1) / (dim+1), summed in vertex order."""
    dim = p.shape[1]
    acc = p[t[:, 0]]
    for k in range(1, dim + 1):
        acc = acc + p[t[:, k]]
    return acc / (dim + 1)


def cull_mask(p, t, fd, geps):
    """mesh_generator.py:734-738 : keep cells whose centroid has fd < -geps."""
    return fd(centroids(p, t)) < -geps


def unique_bars(t):
    """mesh_generator.py:680-688 + geometry/cpp/fast_geometry.cpp:30-77.

    Cell bar pairs ([0,1],[1,2],[2,0] (+[0,3],[1,3],[2,3])), each as (min,max), sorted
    lexicographically, duplicates removed.  Returns int32 (E,2).
    """
    t = np.asarray(t)
    dim = t.shape[1] - 1
    pairs = [(0, 1), (1, 2), (2, 0)]
    if dim == 3:
        pairs += [(0, 3), (1, 3), (2, 3)]
    e = np.concatenate([t[:, list(pr)] for pr in pairs]).astype(np.int64)
    lo = np.minimum(e[:, 0], e[:, 1])
    hi = np.maximum(e[:, 0], e[:, 1])
    key = np.unique((lo << 32) | hi)
    return np.column_stack((key >> 32, key & 0xFFFFFFFF)).astype(np.int32)


def bar_lengths(p, bars):
    """mesh_generator.py:696-698"""
    barvec = p[bars[:, 0]] - p[bars[:, 1]]
    L = np.sqrt((barvec**2).sum(1))
    L[L == 0] = EPS
    return barvec, L


def scatter_forces(bars, Fvec, N):
    """generation/utils.py:48-68 (coo_matrix(...).toarray()) as called at
    mesh_generator.py:706-711: entries are accumulated sequentially in bar order, +Fvec on
    bars[:,0] then -Fvec on bars[:,1]."""
    dim = Fvec.shape[1]
    Ftot = np.zeros((N, dim))
    idx = np.column_stack((bars[:, 0], bars[:, 1])).ravel()
    for k in range(dim):
        val = np.column_stack((Fvec[:, k], -Fvec[:, k])).ravel()
        np.add.at(Ftot[:, k], idx, val)
    return Ftot


def compute_forces(p, t, fh, L0mult, return_parts=False, n_rows=None):
    """mesh_generator.py:691-712.  n_rows (not in the reference; restates dm_plan_set_rows for the
    multi-GPU slabs): only the bars with an end among the first n_rows vertices are kept, so the rows
    >= n_rows of the result are meaningless (ghost copies, overwritten by the halo exchange)."""
    dim = p.shape[1]
    bars = unique_bars(t)
    if n_rows is not None:
        bars = bars[bars[:, 0] < n_rows]
    barvec, L = bar_lengths(p, bars)
    hbars = fh(p[bars].sum(1) / 2)
    scale = ((L**dim).sum() / (hbars**dim).sum()) ** (1.0 / dim)
    L0 = hbars * L0mult * scale
    F = L0 - L
    F[F < 0] = 0
    Fvec = (F / L)[:, None] * barvec
    Ftot = scatter_forces(bars, Fvec, p.shape[0])
    if return_parts:
        return Ftot, dict(bars=bars, L=L, h=hbars, scale=scale)
    return Ftot


def project_points_back_newton(p, fd, deps, hmin, idx):
    """mesh_generator.py:762-784 (one forward-difference Newton step for escaped points)."""
    p = p.copy()
    dim = p.shape[1]
    d = fd(p)
    ix = d > 0.0 if idx == 0 else np.logical_and(d > 0.0, d < hmin / 1.5)
    if ix.any():
        grads = []
        for i in range(dim):
            dv = np.zeros(dim)
            dv[i] = deps
            grads.append((fd(p[ix] + dv) - d[ix]) / deps)
        g2 = sum(g**2 for g in grads)
        g2 = np.where(g2 < deps, deps, g2)
        p[ix] -= (d[ix] * np.vstack(grads) / g2).T
    return p


def force_iteration(p, t, levels, fh, h0, geps, deps, delta_t=0.30, ifix=(), n_rows=None):
    """One pass of the body of the reference's while-loop AFTER the Delaunay step
    (mesh_generator.py:482, 497-521): cull, forces, pfix, update, projection, maxdp."""
    dim = p.shape[1]
    L0mult = 1 + 0.4 / 2 ** (dim - 1)
    keep = cull_mask(p, t, levels[0], geps)
    tk = t[keep]
    Ftot, parts = compute_forces(p, tk, fh, L0mult, return_parts=True, n_rows=n_rows)
    Ftot[list(ifix)] = 0
    pn = p + delta_t * Ftot
    for idx, lv in enumerate(levels):
        pn = project_points_back_newton(pn, lv, deps, h0, idx)
    maxdp = delta_t * np.sqrt((Ftot**2).sum(1)).max()
    return dict(p=pn, keep=keep, Ftot=Ftot, maxdp=maxdp, **parts)


def improve_level_set_newton(p, bid, fd, deps):
    """mesh_generator.py:741-759 given the boundary vertex ids `bid`."""
    p = p.copy()
    dim = p.shape[1]
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


# ----------------------------------------------------------------------------------------
# sliver removal building blocks
# ----------------------------------------------------------------------------------------

_DH_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))


def dihedral_angles(p, t):
    """geometry/cpp/fast_geometry.cpp:351-418 -> (6T,) radians, cell-major."""
    T = len(t)
    out = np.empty((T, 6))
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(6):
            i0 = t[:, _DH_EDGES[i][0]]
            i1 = t[:, _DH_EDGES[i][1]]
            i2 = t[:, _DH_EDGES[5 - i][0]]
            i3 = t[:, _DH_EDGES[5 - i][1]]
            p0 = p[i0]
            v = [p[i1] - p0, p[i2] - p0, p[i3] - p0]
            for k in range(3):
                nrm = np.sqrt(v[k][:, 0] * v[k][:, 0] + v[k][:, 1] * v[k][:, 1] + v[k][:, 2] * v[k][:, 2])
                v[k] = v[k] / nrm[:, None]

            def dot(a, b):
                return (0.0 + a[:, 0] * b[:, 0]) + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]

            def cross(a, b):
                return np.column_stack(
                    (
                        a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                        a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                        a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0],
                    )
                )

            def l2(a):
                return np.sqrt((0.0 + a[:, 0] * a[:, 0]) + a[:, 1] * a[:, 1] + a[:, 2] * a[:, 2])

            v1, v2, v3 = v
            cphi = (dot(v2, v3) - dot(v1, v2) * dot(v1, v3)) / (l2(cross(v1, v2)) * l2(cross(v1, v3)))
            out[:, i] = np.arccos(cphi)
    return out.ravel()


def sliver_cells(p, t, min_dh, max_dh):
    """mesh_generator.py:532-540 : ids of tets with any dihedral angle outside the bounds."""
    dh = dihedral_angles(p, t)
    bad = np.argwhere((dh < min_dh) | (dh > max_dh))
    return np.unique(np.floor(bad / 6).astype("int"))


def circumsphere_grad(p0, p1, p2, p3):
    """geometry/cpp/fast_geometry.cpp:580-666 : gradient of the circumradius wrt vertex 0."""
    x1, y1, z1 = (p0 - p3).T
    x2, y2, z2 = (p1 - p3).T
    x3, y3, z3 = (p2 - p3).T
    sq1 = x1 * x1 + y1 * y1 + z1 * z1
    sq2 = x2 * x2 + y2 * y2 + z2 * z2
    sq3 = x3 * x3 + y3 * y3 + z3 * z3
    dax = y2 * z3 - y3 * z2
    day = z2 * x3 - x2 * z3
    daz = x2 * y3 - x3 * y2
    dDx_dx = -2.0 * x1 * dax
    dDx_dy = -2.0 * y1 * dax + sq2 * z3 - sq3 * z2
    dDx_dz = -2.0 * z1 * dax - sq2 * y3 + sq3 * y2
    dDy_dx = -2.0 * x1 * day - sq2 * z3 + sq3 * z2
    dDy_dy = -2.0 * y1 * day
    dDy_dz = -2.0 * z1 * day + sq2 * x3 - sq3 * x2
    dDz_dx = -2.0 * x1 * daz + sq2 * y3 - sq3 * y2
    dDz_dy = -2.0 * y1 * daz - sq2 * x3 + sq3 * x2
    dDz_dz = -2.0 * z1 * daz
    a = x1 * dax + y1 * day + z1 * daz
    Dx = -sq1 * dax + y1 * (sq2 * z3 - sq3 * z2) - z1 * (sq2 * y3 - sq3 * y2)
    Dy = -sq1 * day - x1 * (sq2 * z3 - sq3 * z2) + z1 * (sq2 * x3 - sq3 * x2)
    Dz = -sq1 * daz + x1 * (sq2 * y3 - sq3 * y2) - y1 * (sq2 * x3 - sq3 * x2)
    ssq = Dx * Dx + Dy * Dy + Dz * Dz
    with np.errstate(invalid="ignore", divide="ignore"):
        gx = (Dx * dDx_dx + Dy * dDy_dx + Dz * dDz_dx) / (2.0 * a * a) - (dax * ssq) / (2.0 * a * a * a)
        gy = (Dx * dDx_dy + Dy * dDy_dy + Dz * dDz_dy) / (2.0 * a * a) - (day * ssq) / (2.0 * a * a * a)
        gz = (Dx * dDx_dz + Dy * dDy_dz + Dz * dDz_dz) / (2.0 * a * a) - (daz * ssq) / (2.0 * a * a * a)
    return np.column_stack((gx, gy, gz))


def sliver_perturbation(p, t, ele_nums, step, h0):
    """mesh_generator.py:245-274 : move vertex 0 of every sliver along the normalised
    circumsphere gradient (fancy-index `+=`: for a repeated vertex the LAST sliver wins)."""
    p = p.copy()
    s = t[ele_nums]
    g = circumsphere_grad(p[s[:, 0]], p[s[:, 1]], p[s[:, 2]], p[s[:, 3]])
    g[np.isinf(g)] = 1.0
    g /= (np.sum(np.abs(g) ** 2, axis=-1) ** 0.5)[:, None]
    p[s[:, 0]] += step * h0 * g
    return p


# ----------------------------------------------------------------------------------------
# initial points
# ----------------------------------------------------------------------------------------


def staggered_grid(h0, dim, bbox):
    """generation/utils.py:15-25 ; bbox is (dim,2)."""
    pts = np.mgrid[tuple(slice(lo, hi + h0, h0) for lo, hi in bbox)].astype(float)
    odd0 = [i for i in range(pts[0].shape[0]) if i % 2]
    odd1 = [i for i in range(pts[1].shape[0]) if i % 2]
    pts[1][odd0] += h0 / 2
    if dim == 3:
        pts[2][odd1] += h0 / 2
    return pts.reshape(dim, -1).T


def initial_points(h0, geps, dim, bbox, fh, fd, pfix, seed=0):
    """mesh_generator.py:808-852 (serial branch, r0m = min(r0))."""
    p = staggered_grid(h0, dim, bbox)
    p = p[fd(p) < geps]
    r0 = fh(p)
    r0m = r0.min()
    np.random.seed(seed)
    return np.vstack((pfix, p[np.random.rand(p.shape[0]) < r0m**dim / r0**dim]))


# ----------------------------------------------------------------------------------------
# quality metrics used by the end-to-end parity checks (geometry/utils.py:175-199, 252-274)
# ----------------------------------------------------------------------------------------


def simp_vol(p, t):
    if p.shape[1] == 2:
        d01 = p[t[:, 1]] - p[t[:, 0]]
        d02 = p[t[:, 2]] - p[t[:, 0]]
        return (d01[:, 0] * d02[:, 1] - d01[:, 1] * d02[:, 0]) / 2
    d01 = p[t[:, 1]] - p[t[:, 0]]
    d02 = p[t[:, 2]] - p[t[:, 0]]
    d03 = p[t[:, 3]] - p[t[:, 0]]
    return np.einsum("ij,ij->i", np.cross(d01, d02), d03) / 6


def simp_qual(p, t):
    def length(v):
        return np.sqrt((v**2).sum(1))

    a = length(p[t[:, 1]] - p[t[:, 0]])
    b = length(p[t[:, 2]] - p[t[:, 0]])
    c = length(p[t[:, 2]] - p[t[:, 1]])
    r = 0.5 * np.sqrt((b + c - a) * (c + a - b) * (a + b - c) / (a + b + c))
    R = a * b * c / np.sqrt((a + b + c) * (b + c - a) * (c + a - b) * (a + b - c))
    return 2 * r / R


__all__ = [n for n in dir() if not n.startswith("_")]
_ = math


# ----------------------------------------------------------------------------------------
# Sizing preprocessing (SURVEY section 8f "next #3")
# ----------------------------------------------------------------------------------------
def limgrad(f, delta, ftol, max_sweeps=100000):
    """Fixed point of the reference's gradient limiter (sizing/cpp/FastHJ.cpp:63-157): every pair
    of stencil neighbours (6 clamped edges) is relaxed until no node exceeds a neighbour by more
    than delta (+ftol).  The reference visits pairs in a sequential active-set order; the operator
    is monotone (values only go down) so the fixed point does not depend on that order (up to
    ftol) -- restated here as vectorised sweeps.  Pinned against the reference's compiled _FastHJ
    through tests/golden/sizing_*.npz."""
    f = np.array(f, dtype=np.float64, copy=True)
    nd = f.ndim
    for _ in range(max_sweeps):
        m = f.copy()
        for ax in range(nd):
            lo = [slice(None)] * nd
            hi = [slice(None)] * nd
            lo[ax], hi[ax] = slice(0, -1), slice(1, None)
            m[tuple(hi)] = np.minimum(m[tuple(hi)], f[tuple(lo)])
            m[tuple(lo)] = np.minimum(m[tuple(lo)], f[tuple(hi)])
        cand = m + delta
        upd = f > cand + ftol
        if not upd.any():
            break
        f[upd] = cand[upd]
    return f


def sizing_from_velocity(vp, bbox, hmin=150.0, hmax=10000.0, wl=0, freq=2.0, grad=0.0, grade=0.0, stencil_size=10.0,
                         space_order=1, dt=0.0, cr_max=1.0, pad_style="edge", domain_pad=0.0, **_unused):
    """get_sizing_function_from_segy for a velocity array (sizing/mesh_size_function.py:128-232):
    returns (cell_size grid, padded bbox).  Step order as in the reference: wavelength / gradient
    sizing (:411-450), hmin / hmax clamps (:214-218), CFL limit (:453-468), gradation (:471-496),
    domain pad (:526-587)."""
    from scipy import ndimage

    vp = np.array(vp, dtype=np.float64, copy=True)
    dim = vp.ndim
    nz = vp.shape[0]
    cell = np.full(vp.shape, hmin, dtype=float)
    if wl > 0 or grad > 0:
        h_wl = 99999 if wl == 0.0 else vp / (freq * wl)
        h_gr = 99999
        if grad != 0.0:
            window = [stencil_size] * dim if np.isscalar(stencil_size) else stencil_size
            mean = ndimage.uniform_filter(vp, tuple(window))
            var = ndimage.uniform_filter(vp**2, tuple(window)) - mean**2
            var = np.divide(var, np.amax(var))
            var -= np.amin(var)
            h_gr = grad / (var + 0.10)
        cell = np.minimum(h_wl, h_gr)
    cell[cell < hmin] = hmin
    cell[cell > hmax] = hmax
    if not (cr_max == 0.0 or dt == 0.0 or space_order == 0.0):
        cr_old = (vp * dt) / (dim * cell)
        lim = cr_max / (dim * space_order)
        cell = np.where(cr_old > lim, (vp * dt) / (dim * lim), cell)
    if grade != 0.0:
        elen = (bbox[1] - bbox[0]) / nz
        cell = limgrad(cell, elen * grade, cell.min() * math.sqrt(1e-9))
    if domain_pad > 0:
        d = [(bbox[2 * k + 1] - bbox[2 * k]) / vp.shape[k] for k in range(dim)]
        nn = [int(domain_pad / dk) for dk in d]
        bbox = tuple(v for k in range(dim)
                     for v in (bbox[2 * k] - domain_pad, bbox[2 * k + 1] + (domain_pad if k > 0 else 0.0)))
        padding = tuple((nn[k], 0) if k == 0 else (nn[k], nn[k]) for k in range(dim))
        mx = np.amax(cell)
        if pad_style == "edge":
            cell = np.pad(cell, padding, "edge")
        elif pad_style == "constant":
            cell = np.pad(cell, padding, "constant", constant_values=(mx, mx))
        elif pad_style == "linear_ramp":
            cell = np.pad(cell, padding, "linear_ramp", end_values=(mx, mx))
        else:
            raise ValueError("pad style currently not supported. Try `linear_ramp`, `edge`, or `constant`")
    return cell, tuple(bbox)


def cells_lead_interior(t, key, thresh):
    """Column order of the tetrahedra handed to the sliver loop (restates dm_cells_lead_interior,
    include/distmesh_b200.h; not a reference function: the reference moves `t[ele, 0]`,
    mesh_generator.py:234,245-274, with whatever order CGAL's get_finite_cells has).  Column 0 becomes,
    among the vertices with key < thresh, the one picked by (sum of the ids) mod (their count), and the
    vertex with the smallest key when there is none; applied as an even permutation."""
    t = np.asarray(t)
    kk = np.asarray(key)[t]
    ok = kk < thresh
    cnt = ok.sum(axis=1)
    pick = np.where(cnt > 0, t.astype(np.int64).sum(axis=1) % np.maximum(cnt, 1), 0)
    cs = np.cumsum(ok, axis=1)
    sel = np.argmax((cs == (pick[:, None] + 1)) & ok, axis=1)
    sel = np.where(cnt > 0, sel, np.argmin(kk, axis=1))
    even = np.array([[0, 1, 2, 3], [1, 2, 0, 3], [2, 0, 1, 3], [3, 1, 0, 2]])
    return np.ascontiguousarray(np.take_along_axis(t, even[sel], axis=1))


def pad_staged(a, padding, mode, end_values):
    """np.pad(a, padding, mode) for mode in ("edge", "constant", "linear_ramp"), restated the way
    numpy/lib/_arraypad_impl.py works (NumPy 2.3; the reference calls it at sizing/mesh_size_function.py:575-587):
    one axis after the other, axis 0 first; while axis k is padded the region of interest is the full (already
    padded) extent along the axes before k and the original extent along the axes after k (_view_roi); a linear
    ramp is np.linspace(end_value, edge, width, endpoint=False), i.e. j * ((edge - end) / width) + end, or --
    when ANY element of the edge plane has a zero step -- (j / width) * (edge - end) + end for the whole plane
    (numpy/_core/function_base.py: any_step_zero).  The CUDA kernel dm_pad follows this restatement."""
    a = np.asarray(a, dtype=np.float64)
    nd = a.ndim
    sl = tuple(slice(padding[k][0], padding[k][0] + a.shape[k]) for k in range(nd))
    P = np.empty([a.shape[k] + padding[k][0] + padding[k][1] for k in range(nd)])
    P[sl] = a
    for k in range(nd):
        roi = P[tuple(slice(None) if j <= k else sl[j] for j in range(nd))]
        lw, rw = padding[k]
        n = roi.shape[k]
        for lower, W, ev in ((True, lw, end_values[0]), (False, rw, end_values[1])):
            if W == 0:
                continue
            ix = [slice(None)] * nd
            ix[k] = slice(lw, lw + 1) if lower else slice(n - rw - 1, n - rw)
            edge = roi[tuple(ix)]
            if mode == "edge":
                y = np.repeat(edge, W, axis=k)
            elif mode == "constant":
                y = np.full_like(np.repeat(edge, W, axis=k), ev)
            else:
                delta = edge - ev
                step = delta / W
                j = np.arange(W, dtype=np.float64).reshape([-1 if q == k else 1 for q in range(nd)])
                y = (j / W) * delta if (step == 0).any() else j * step
                y = y + ev
            ix[k] = slice(0, W) if lower else slice(n - W, n)
            roi[tuple(ix)] = y if lower else np.flip(y, axis=k)
    return P


def uniform_filter1d_lines(a, size, axis):
    """scipy.ndimage.uniform_filter1d(a, size, axis=axis) (mode "reflect", origin 0; SciPy 1.18, C source
    ni_filters.c NI_UniformFilter1D), restated: per line a running sum over the reflect-extended samples --
    tmp = sum(ext[0:size]); out[0] = tmp / size; tmp += ext[i + size - 1] - ext[i - 1]; out[i] = tmp / size.
    scipy.ndimage.uniform_filter applies it axis after axis (axis 0 first); the reference calls that at
    sizing/mesh_size_function.py:436-437.  The CUDA kernel dm_uniform_filter follows this restatement."""
    a = np.moveaxis(np.asarray(a, dtype=np.float64), axis, -1)
    n = a.shape[-1]
    left = size // 2
    idx = np.arange(n + size - 1) - left
    idx = np.where(idx < 0, -idx - 1, idx)
    idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
    ext = a[..., idx]
    out = np.empty_like(a)
    tmp = np.zeros(a.shape[:-1])
    for k in range(size):
        tmp = tmp + ext[..., k]
    out[..., 0] = tmp / size
    for i in range(1, n):
        tmp = tmp + (ext[..., i + size - 1] - ext[..., i - 1])
        out[..., i] = tmp / size
    return np.moveaxis(out, -1, axis)


def uniform_filter(a, size):
    out = np.asarray(a, dtype=np.float64)
    for axis, s in enumerate(size):
        out = uniform_filter1d_lines(out, int(s), axis)
    return out
