"""Host Delaunay triangulator (libdistmesh_host.so, include/distmesh_host.h) -- the retriangulation
step north_star keeps on the host (reference: CGAL behind generation/cpp/delaunay_class.cpp).

Checked on the CPU: exact predicates against rational arithmetic, the cell set against Qhull on
points in general position, and the Delaunay / partition properties on the degenerate inputs DistMesh
produces (the initial staggered lattice, co-circular points, boundary vertices on straight edges,
duplicates, collinear input).  Parity with the reference's CGAL is unpinned by construction (CGAL is
not installed); any correct Delaunay code gives the same cells for points in general position."""
import ctypes as C
import os
import re
from fractions import Fraction

import numpy as np
import pytest
from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "distmesh_host.h")


@pytest.fixture(scope="module")
def hl():
    import subprocess

    from seismicmesh_b200 import _hostlib

    host = os.path.join(ROOT, "seismicmesh_b200", "csrc", "host")
    newest = max(os.path.getmtime(os.path.join(host, f)) for f in os.listdir(host))
    if not os.path.exists(_hostlib.LIB_PATH) or os.path.getmtime(_hostlib.LIB_PATH) < newest:
        subprocess.check_call(["bash", os.path.join(host, "build.sh")])

    return _hostlib.lib()


@pytest.fixture(scope="module")
def tri(hl):
    from seismicmesh_b200.triangulator import SweepHullTriangulator

    return SweepHullTriangulator(2)


def _canon(t):
    t = np.sort(np.asarray(t, dtype=np.int64), axis=1)
    return t[np.lexsort(t.T[::-1])]


def _signed_area(p, t):
    a = p[t]
    return 0.5 * ((a[:, 1, 0] - a[:, 0, 0]) * (a[:, 2, 1] - a[:, 0, 1]) - (a[:, 1, 1] - a[:, 0, 1]) * (a[:, 2, 0] - a[:, 0, 0]))


def _sgn(x):
    return (x > 0) - (x < 0)


def _is_lex_sorted(t):
    """The documented output order: cells in lexicographic order of their SORTED ids (hence grouped by
    their smallest id), each cell keeping the triangulator's own column order."""
    t = np.asarray(t, dtype=np.int64)
    if len(t) == 0:
        return True
    ts = np.sort(t, axis=1)
    if not (np.diff(ts, axis=1) > 0).all():
        return False
    return np.array_equal(ts, ts[np.lexsort(ts.T[::-1])])


def _column0_unbiased(t):
    """Column 0 must not be systematically the smallest (or largest) id of its cell: the reference's
    sliver_removal moves t[:,0] of every sliver (mesh_generator.py:245-274)."""
    t = np.asarray(t, dtype=np.int64)
    if len(t) < 200:
        return True
    frac_min = (t[:, 0] == t.min(axis=1)).mean()
    frac_max = (t[:, 0] == t.max(axis=1)).mean()
    return frac_min < 0.6 and frac_max < 0.6


def _incircle(hl, p, i, j, k, m):
    q = [np.ascontiguousarray(p[x], dtype=np.float64) for x in (i, j, k, m)]
    return hl.dmh_incircle(*[x.ctypes.data for x in q])


def _check_triangulation(hl, p, t, n_used=None):
    """ccw, non-degenerate, manifold, covers the convex hull, locally (hence globally) Delaunay."""
    from scipy.spatial import ConvexHull

    area = np.abs(_signed_area(p, t))
    assert (area > 0).all()
    assert _is_lex_sorted(t)
    hull = ConvexHull(p).volume
    assert abs(area.sum() - hull) <= 1e-9 * hull
    assert np.unique(t).size == (len(p) if n_used is None else n_used)
    edges = {}
    for ti, (i, j, k) in enumerate(t.tolist()):
        for u, v, w in ((i, j, k), (j, k, i), (k, i, j)):
            edges.setdefault((min(u, v), max(u, v)), []).append((ti, w))
    for lst in edges.values():
        assert len(lst) <= 2
        if len(lst) == 2:
            (t1, _), (_, w2) = lst
            i, j, k = t[t1]
            if _signed_area(p, t[t1:t1 + 1])[0] < 0:
                j, k = k, j  # incircle wants a counter-clockwise triangle; the output order is by id
            assert _incircle(hl, p, i, j, k, w2) <= 0  # the opposite vertex is not strictly inside


def test_library_exports_every_declared_symbol(hl):
    from seismicmesh_b200 import _hostlib

    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(dmh_[a-z0-9_]+)\s*\(", src)))
    assert declared == sorted(_hostlib.EXPORTED_SYMBOLS)
    for n in declared:
        assert hasattr(hl, n)
    assert hl.dmh_version().startswith(b"distmesh_host")


def test_predicates_are_exact(hl):
    """Sign of orient2d / incircle against rational arithmetic on random, lattice, nearly co-circular
    and nearly collinear quadruples (the filter must hand the close calls to the exact path)."""
    rng = np.random.default_rng(11)

    def o2(a, b, c):
        a, b, c = [[Fraction(float(v)) for v in q] for q in (a, b, c)]
        return _sgn((a[0] - c[0]) * (b[1] - c[1]) - (a[1] - c[1]) * (b[0] - c[0]))

    def ic(a, b, c, d):
        a, b, c, d = [[Fraction(float(v)) for v in q] for q in (a, b, c, d)]
        r = []
        for q in (a, b, c):
            x, y = q[0] - d[0], q[1] - d[1]
            r.append((x, y, x * x + y * y))
        (a0, a1, a2), (b0, b1, b2), (c0, c1, c2) = r
        return _sgn(a0 * (b1 * c2 - b2 * c1) - a1 * (b0 * c2 - b2 * c0) + a2 * (b0 * c1 - b1 * c0))

    zeros = 0
    for trial in range(1500):
        kind = trial % 4
        if kind == 0:
            pts = rng.random((4, 2))
        elif kind == 1:  # lattice points far from the origin: exact ties
            pts = rng.integers(0, 5, (4, 2)).astype(float) * 0.1 + rng.integers(0, 2) * 1e3
        elif kind == 2:  # on a circle up to rounding
            th = rng.random(4) * 6.28
            pts = rng.random(2) + np.c_[np.cos(th), np.sin(th)] * 0.5
        else:  # three points on a line up to rounding
            a, d = rng.random(2), rng.random(2)
            pts = np.array([a + d * s for s in rng.random(4)])
            pts[3] = rng.random(2)
        q = [np.ascontiguousarray(x) for x in pts]
        e1, e2 = o2(*q[:3]), ic(*q)
        assert _sgn(hl.dmh_orient2d(*[x.ctypes.data for x in q[:3]])) == e1
        assert _sgn(hl.dmh_incircle(*[x.ctypes.data for x in q])) == e2
        zeros += (e1 == 0) + (e2 == 0)
    assert zeros > 50  # the exact-tie path was exercised


@pytest.mark.parametrize("n,seed", [(3, 0), (4, 1), (10, 2), (100, 3), (5000, 4), (60000, 5)])
def test_same_cells_as_qhull_in_general_position(tri, n, seed):
    from scipy.spatial import Delaunay

    p = np.random.default_rng(seed).random((n, 2)) * [3.0, 1.0] - [1.0, 0.5]
    t = tri.triangulate(p)
    assert t.dtype == np.int32 and t.flags.c_contiguous and t.shape[1] == 2 + 1
    assert (_signed_area(p, t) != 0).all() and _is_lex_sorted(t)
    assert np.array_equal(_canon(t), _canon(Delaunay(p).simplices))
    assert tri.qhull_retries == 0


def test_distmesh_iterate_same_cells_as_qhull(tri):
    """A DistMesh-shaped input: the reference's staggered lattice inside the unit disk, jittered."""
    from scipy.spatial import Delaunay

    from seismicmesh_b200.generation import _staggered_grid

    h0 = 0.02
    p = _staggered_grid(h0, 2, np.array([[-1.0, 1.0], [-1.0, 1.0]]))
    p = p[np.sqrt((p**2).sum(1)) - 1.0 < 0.1 * h0]
    p = np.ascontiguousarray(p + np.random.default_rng(0).uniform(-0.1 * h0, 0.1 * h0, p.shape))
    assert np.array_equal(_canon(tri.triangulate(p)), _canon(Delaunay(p).simplices))


def test_degenerate_inputs(hl, tri):
    rng = np.random.default_rng(7)
    # square lattice: every cell pair is co-circular (flips must terminate; any diagonal is Delaunay)
    g = np.stack(np.meshgrid(np.arange(30.0), np.arange(40.0), indexing="ij"), -1).reshape(-1, 2) * 0.1
    t = tri.triangulate(g)
    assert len(t) == 2 * 29 * 39
    _check_triangulation(hl, g, t)
    # the reference's initial lattice (generation/utils.py:15-25), un-jittered
    from seismicmesh_b200.generation import _staggered_grid

    s = np.ascontiguousarray(_staggered_grid(0.05, 2, np.array([[-1.0, 1.0], [-1.0, 1.0]])))
    _check_triangulation(hl, s, tri.triangulate(s))
    # vertices projected onto the straight edges of a rectangle: collinear hull vertices stay in
    b = rng.random((500, 2))
    b[:100, 0], b[100:200, 0], b[200:300, 1], b[300:400, 1] = 0.0, 1.0, 0.0, 1.0
    t = tri.triangulate(b)
    _check_triangulation(hl, b, t)
    assert len(t) == 2 * 500 - 2 - 400  # all 400 boundary vertices are on the hull
    # co-circular points around a centre
    th = np.linspace(0, 2 * np.pi, 64, endpoint=False)
    c = np.r_[np.c_[np.cos(th), np.sin(th)], [[0.0, 0.0]]]
    t = tri.triangulate(c)
    assert len(t) == 64
    _check_triangulation(hl, c, t)
    # exact duplicates: the first copy is kept, the others are in no cell (as with Qhull)
    d = rng.random((200, 2))
    d2 = np.r_[d, d[:50]]
    t = tri.triangulate(d2)
    assert t.max() < 200
    _check_triangulation(hl, d, t)
    assert tri.qhull_retries == 0


def test_random_degenerate_stress_2d(hl, tri):
    """Massively degenerate random inputs (tiny integer lattices full of duplicates, all points on a
    circle, half of them on a line): every distinct point used, hull filled, no circumcircle contains
    a vertex (exact predicates), without Qhull."""
    rng = np.random.default_rng(321)
    done = 0
    for trial in range(120):
        kind, n = trial % 4, int(rng.integers(3, 150))
        if kind == 0:
            p = rng.integers(0, 5, (n, 2)).astype(float)
        elif kind == 1:
            p = rng.integers(0, 9, (n, 2)).astype(float) * 0.1 - 3.0
        elif kind == 2:
            th = rng.random(n) * 6.283
            p = np.c_[np.cos(th), np.sin(th)]
        else:
            p = rng.random((n, 2))
            p[: n // 2, 1] = 0.25
        p = np.ascontiguousarray(p)
        t = tri.triangulate(p)
        if len(t) == 0:
            continue
        _, first, inv = np.unique(p, axis=0, return_index=True, return_inverse=True)
        tm = np.sort(first[np.asarray(inv).ravel()][t], axis=1)
        tm = tm[np.lexsort(tm.T[::-1])].astype(np.int32)
        _check_triangulation(hl, p, tm, n_used=len(np.unique(p, axis=0)))
        done += 1
    assert done >= 100 and tri.qhull_retries == 0


def test_empty_collinear_and_coincident_inputs(hl, tri):
    assert tri.triangulate(np.zeros((0, 2))).shape == (0, 3)
    assert tri.triangulate(np.array([[0.0, 0.0], [1.0, 0.0]])).shape == (0, 3)
    assert tri.triangulate(np.zeros((5, 2))).shape == (0, 3)
    line = np.c_[np.arange(10.0), 2 * np.arange(10.0)]
    assert tri.triangulate(line).shape == (0, 3)
    one = tri.triangulate(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]))
    assert one.shape == (1, 3) and sorted(one[0].tolist()) == [0, 1, 2]
    # C ABI error codes: capacity too small, null output
    p = np.random.default_rng(0).random((50, 2))
    out = np.empty((4, 3), np.int32)
    T, dups, lost = C.c_int64(), C.c_int64(), C.c_int64()
    assert hl.dmh_delaunay2d(p.ctypes.data, 50, out.ctypes.data, 4, C.byref(T), C.byref(dups), C.byref(lost)) == -2
    assert T.value > 4  # the required size is reported
    assert hl.dmh_delaunay2d(p.ctypes.data, 50, None, 4, C.byref(T), None, None) == -1
    assert hl.dmh_delaunay2d(p.ctypes.data, -1, out.ctypes.data, 4, C.byref(T), None, None) == -1
    bad = p.copy()
    bad[3, 0] = np.inf
    big0 = np.empty((hl.dmh_delaunay2d_max_cells(50), 3), np.int32)
    assert hl.dmh_delaunay2d(bad.ctypes.data, 50, big0.ctypes.data, len(big0), C.byref(T), None, None) == -1
    # the two reasons a row can be left out are reported separately
    d = np.r_[p, p[:7]]
    big = np.empty((hl.dmh_delaunay2d_max_cells(57), 3), np.int32)
    assert hl.dmh_delaunay2d(d.ctypes.data, 57, big.ctypes.data, len(big), C.byref(T), C.byref(dups), C.byref(lost)) == 0
    assert (dups.value, lost.value) == (7, 0)
    line = np.ascontiguousarray(np.c_[np.arange(10.0), np.arange(10.0)])
    assert hl.dmh_delaunay2d(line.ctypes.data, 10, big.ctypes.data, len(big), C.byref(T), C.byref(dups), C.byref(lost)) == 0
    assert (T.value, dups.value, lost.value) == (0, 0, 10)


def test_get_triangulator_defaults():
    from seismicmesh_b200.triangulator import QhullTriangulator, SweepHullTriangulator, get_triangulator

    assert isinstance(get_triangulator(None, 2), SweepHullTriangulator)
    from seismicmesh_b200.triangulator import BowyerWatsonTriangulator

    assert isinstance(get_triangulator(None, 3), BowyerWatsonTriangulator)
    assert isinstance(get_triangulator("qhull", 2), QhullTriangulator)
    assert isinstance(get_triangulator("qhull", 3), QhullTriangulator)
    assert isinstance(get_triangulator("native", 2), SweepHullTriangulator)
    assert isinstance(get_triangulator("native", 3), BowyerWatsonTriangulator)
    with pytest.raises(ValueError):
        get_triangulator("sweephull", 3)
    with pytest.raises(ValueError):
        get_triangulator("bowyer-watson", 2)
    with pytest.raises(ValueError):
        get_triangulator("cgal", 2)


@pytest.mark.parametrize("k", [2, 3, 4])
def test_sort_unique_rows_matches_numpy(hl, k):
    """dmh_sort_unique_rows_i32 (the termination path's unique_rows of row-wise sorted cells / facets /
    edges): rows, order and multiplicities of np.unique(np.sort(rows, 1), axis=0, return_counts=True),
    below and above the size where it goes multi-threaded; bad ids are refused."""
    rng = np.random.default_rng(k)
    for n, N in ((0, 5), (1, 3), (500, 40), (150000, 3000)):
        rows = rng.integers(0, N, (n, k)).astype(np.int32)
        rows[: n // 3] = rows[n // 3: 2 * (n // 3)]  # plenty of duplicates
        want, wc = (np.unique(np.sort(rows, axis=1), axis=0, return_counts=True) if n else (rows, np.empty(0, int)))
        got = np.ascontiguousarray(rows.copy())
        counts = np.empty(max(n, 1), np.int32)
        nu = C.c_int64(-1)
        assert hl.dmh_sort_unique_rows_i32(got.ctypes.data, n, k, N, counts.ctypes.data, C.byref(nu), 4) == 0
        assert nu.value == len(want) and np.array_equal(got[: nu.value], want) and np.array_equal(counts[: nu.value], wc)
        nu2 = C.c_int64(-1)
        again = np.ascontiguousarray(rows.copy())
        assert hl.dmh_sort_unique_rows_i32(again.ctypes.data, n, k, N, None, C.byref(nu2), 1) == 0
        assert nu2.value == len(want) and np.array_equal(again[: nu2.value], want)
    bad = np.array([[0, 1, 7][:k] + [0] * (k - 3)], dtype=np.int32)
    bad[0, -1] = 9
    assert hl.dmh_sort_unique_rows_i32(bad.ctypes.data, 1, k, 9, None, C.byref(nu), 1) == -1
    assert hl.dmh_sort_unique_rows_i32(bad.ctypes.data, 1, 5, 10, None, C.byref(nu), 1) == -1
