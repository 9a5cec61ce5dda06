"""Host 3-D Delaunay triangulator (dmh_delaunay3d in libdistmesh_host.so) -- the retriangulation step
north_star keeps on the host (reference: CGAL behind generation/cpp/delaunay_class3.cpp).

Checked on the CPU: exact predicates against rational arithmetic, the cell set against Qhull on points
in general position (random and DistMesh iterates), and orientation / partition / empty-circumsphere
properties on the degenerate inputs DistMesh produces (cubic and the reference's staggered lattice,
vertices on a sphere and on the faces of a cube, duplicates, coplanar input)."""
import ctypes as C
from fractions import Fraction

import numpy as np
import pytest

from test_host_delaunay import _column0_unbiased, _is_lex_sorted, hl  # noqa: F401  (fixture: builds / loads libdistmesh_host.so)


@pytest.fixture(scope="module")
def tri(hl):  # noqa: F811
    from seismicmesh_b200.triangulator import BowyerWatsonTriangulator

    return BowyerWatsonTriangulator(3)


def _canon(t):
    t = np.sort(np.asarray(t, dtype=np.int64), axis=1)
    return t[np.lexsort(t.T[::-1])]


def _sgn(x):
    return (x > 0) - (x < 0)


def _det(m):
    if len(m) == 1:
        return m[0][0]
    return sum((-1) ** j * m[0][j] * _det([r[:j] + r[j + 1:] for r in m[1:]]) for j in range(len(m)))


def _ptrs(*pts):
    keep = [np.ascontiguousarray(x, dtype=np.float64) for x in pts]
    return keep, [x.ctypes.data for x in keep]


def _check_triangulation(hl, p, t, n_used=None):  # noqa: F811
    """positively oriented, non-degenerate, manifold, fills the convex hull, empty circumspheres."""
    from scipy.spatial import ConvexHull

    a = p[t]
    vol = np.einsum("ij,ij->i", np.cross(a[:, 1] - a[:, 0], a[:, 2] - a[:, 0]), a[:, 3] - a[:, 0]) / 6
    # every cell is non-degenerate by the EXACT predicate (a sliver of four nearly coplanar hull
    # vertices may have a volume of 1e-19 that rounds to anything in floating point)
    orient = []
    for tt in t:
        keep, q = _ptrs(*p[tt])
        orient.append(hl.dmh_orient3d(*q))
        assert orient[-1] != 0
    assert _is_lex_sorted(t)
    hull = ConvexHull(p).volume
    assert abs(np.abs(vol).sum() - hull) <= 1e-9 * hull
    assert np.unique(t).size == (len(p) if n_used is None else n_used)
    faces = {}
    for ti, tt in enumerate(t.tolist()):
        for k in range(4):
            faces.setdefault(tuple(sorted(tt[:k] + tt[k + 1:])), []).append((ti, tt[k]))
    for lst in faces.values():
        assert len(lst) <= 2
        if len(lst) == 2:
            (t1, _), (_, w2) = lst
            keep, q = _ptrs(*[p[x] for x in t[t1]], p[w2])
            # insphere's sign means "inside" for a positively oriented cell; the output order is by id
            assert hl.dmh_insphere(*q) * orient[t1] <= 0  # the opposite vertex is not strictly inside


def test_predicates_are_exact(hl):  # noqa: F811
    rng = np.random.default_rng(1)

    def o3(a, b, c, d):
        q = [[Fraction(float(v)) for v in x] for x in (a, b, c, d)]
        return _sgn(_det([[q[i][k] - q[3][k] for k in range(3)] for i in range(3)]))

    def isph(a, b, c, d, e):
        q = [[Fraction(float(v)) for v in x] for x in (a, b, c, d, e)]
        rows = []
        for i in range(4):
            r = [q[i][k] - q[4][k] for k in range(3)]
            rows.append(r + [sum(v * v for v in r)])
        return _sgn(_det(rows))

    # convention: for a positively oriented tetrahedron the centroid is inside (insphere > 0)
    tet = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    keep, q = _ptrs(*tet)
    if hl.dmh_orient3d(*q) < 0:
        tet = tet[[1, 0, 2, 3]]
    keep, q = _ptrs(*tet, tet.mean(0))
    assert hl.dmh_orient3d(*q[:4]) > 0 and hl.dmh_insphere(*q) > 0
    zeros = 0
    for trial in range(800):
        kind = trial % 4
        if kind == 0:
            pts = rng.random((5, 3))
        elif kind == 1:  # lattice points far from the origin: exact ties
            pts = rng.integers(0, 4, (5, 3)).astype(float) * 0.25 + rng.integers(0, 2) * 100
        elif kind == 2:  # on a sphere up to rounding
            v = rng.normal(size=(5, 3))
            pts = rng.random(3) + 0.5 * v / np.linalg.norm(v, axis=1)[:, None]
        else:  # four points in a plane up to rounding
            o, u, w = rng.random(3), rng.random(3), rng.random(3)
            pts = np.array([o + s * u + t * w for s, t in rng.random((5, 2))])
            pts[4] = rng.random(3)
        keep, q = _ptrs(*pts)
        e1, e2 = o3(*pts[:4]), isph(*pts)
        assert _sgn(hl.dmh_orient3d(*q[:4])) == e1
        assert _sgn(hl.dmh_insphere(*q)) == e2  # the same determinant: its sign means 'inside' when e1 > 0
        zeros += (e1 == 0) + (e2 == 0)
    assert zeros > 40


@pytest.mark.parametrize("n,seed", [(4, 0), (5, 1), (20, 2), (200, 3), (5000, 4), (40000, 5)])
def test_same_cells_as_qhull_in_general_position(tri, n, seed):
    from scipy.spatial import Delaunay

    p = np.random.default_rng(seed).random((n, 3)) * [3.0, 1.0, 0.5] - [1.0, 0.5, 0.0]
    t = tri.triangulate(p)
    assert t.dtype == np.int32 and t.flags.c_contiguous and t.shape[1] == 4 and _is_lex_sorted(t)
    assert _column0_unbiased(t)
    assert np.array_equal(_canon(t), _canon(Delaunay(p).simplices))
    assert tri.qhull_retries == 0


def test_distmesh_iterates(hl, tri):  # noqa: F811
    """Ball, h0 = 0.15, oracle force iterations.  From the reference's own initial points the interior
    stays a lattice (balanced forces), i.e. co-spherical: the triangulation must be a valid Delaunay
    one (ties may be broken differently from Qhull).  From jittered initial points everything is in
    general position and the cells must EQUAL Qhull's, iteration after iteration."""
    from scipy.spatial import Delaunay

    from oracle import distmesh_oracle as orc

    h0 = 0.15
    spec = ("ball", {"x0": [0.0, 0.0, 0.0], "r": 1.0})
    fd = lambda x: orc.sdf(spec, x)  # noqa: E731
    fh = lambda x: np.full(len(x), h0)  # noqa: E731
    geps, deps = 0.1 * h0, np.sqrt(np.finfo(float).eps) * h0
    p0 = orc.initial_points(h0, geps, 3, np.array([[-1.0, 1.0]] * 3), fh, fd, np.empty((0, 3)))
    p = p0.copy()
    for it in range(3):
        t = tri.triangulate(p)
        _check_triangulation(hl, p, t)
        p = orc.force_iteration(p, t, [fd], fh, h0, geps, deps)["p"]
    p = p0 + np.random.default_rng(0).uniform(-0.1 * h0, 0.1 * h0, p0.shape)
    for it in range(4):
        t = tri.triangulate(p)
        assert np.array_equal(_canon(t), _canon(Delaunay(p).simplices))
        p = orc.force_iteration(p, t, [fd], fh, h0, geps, deps)["p"]
    assert tri.qhull_retries == 0


def test_degenerate_inputs(hl, tri):  # noqa: F811
    rng = np.random.default_rng(7)
    g = np.stack(np.meshgrid(np.arange(7.0), np.arange(8.0), np.arange(9.0), indexing="ij"), -1).reshape(-1, 3) * 0.1
    t = tri.triangulate(g)  # cubic lattice: every cube is co-spherical
    assert len(t) >= 5 * 6 * 7 * 8
    _check_triangulation(hl, g, t)
    from seismicmesh_b200.generation import _staggered_grid

    s = np.ascontiguousarray(_staggered_grid(0.25, 3, np.array([[-1.0, 1.0]] * 3)))  # generation/utils.py:15-25
    _check_triangulation(hl, s, tri.triangulate(s))
    b = rng.normal(size=(1500, 3))  # ball: vertices exactly on the sphere (up to rounding) + interior
    b /= np.linalg.norm(b, axis=1)[:, None]
    b[:900] *= rng.random((900, 1)) ** (1 / 3)
    _check_triangulation(hl, b, tri.triangulate(b))
    c = rng.random((1500, 3))  # cube: vertices exactly on the faces
    for k in range(3):
        c[100 * (2 * k): 100 * (2 * k + 1), k] = 0.0
        c[100 * (2 * k + 1): 100 * (2 * k + 2), k] = 1.0
    _check_triangulation(hl, c, tri.triangulate(c))
    d = rng.random((300, 3))  # exact duplicates: one copy of each is in the cells
    d2 = np.r_[d, d[:25]]
    t = tri.triangulate(d2)
    assert _is_lex_sorted(t)
    t = np.where(t >= 300, t - 300, t)
    t = np.sort(t, axis=1)
    _check_triangulation(hl, d, t[np.lexsort(t.T[::-1])])
    assert tri.qhull_retries == 0


def test_degenerate_and_error_paths(hl):  # noqa: F811
    T, dups, lost = C.c_int64(), C.c_int64(), C.c_int64()
    out = np.empty((64, 4), np.int32)
    flat = np.ascontiguousarray(np.c_[np.random.default_rng(0).random((30, 2)), np.zeros(30)])
    assert hl.dmh_delaunay3d(flat.ctypes.data, 30, out.ctypes.data, 64, C.byref(T), C.byref(dups), C.byref(lost)) == 0
    assert (T.value, lost.value) == (0, 30)  # coplanar input: no cell, every row reported
    assert hl.dmh_delaunay3d(flat.ctypes.data, 3, out.ctypes.data, 64, C.byref(T), None, None) == 0 and T.value == 0
    p = np.random.default_rng(1).random((100, 3))
    assert hl.dmh_delaunay3d(p.ctypes.data, 100, out.ctypes.data, 64, C.byref(T), None, None) == -2
    assert T.value > 64  # the size that is needed
    assert hl.dmh_delaunay3d(p.ctypes.data, 100, None, 64, C.byref(T), None, None) == -1
    bad = p.copy()
    bad[5, 1] = np.nan
    big = np.empty((hl.dmh_delaunay3d_max_cells(100), 4), np.int32)
    assert hl.dmh_delaunay3d(bad.ctypes.data, 100, big.ctypes.data, len(big), C.byref(T), None, None) == -1


def _first_occurrence_cells(p, t):
    """duplicate rows -> their first occurrence, ids ascending, cells lex-sorted"""
    _, first, inv = np.unique(p, axis=0, return_index=True, return_inverse=True)
    tm = np.sort(first[np.asarray(inv).ravel()][t], axis=1)
    return tm[np.lexsort(tm.T[::-1])].astype(np.int32)


def test_random_degenerate_stress_3d(hl, tri):  # noqa: F811
    """Massively degenerate random inputs: tiny integer lattices full of duplicates, all points on a
    sphere, half of them in a plane, a collinear bunch.  Every distinct point must be used, the cells
    must fill the hull, and no circumball may contain a vertex (exact predicates), without Qhull."""
    rng = np.random.default_rng(123)
    done = 0
    for trial in range(100):
        kind, n = trial % 5, int(rng.integers(5, 100))
        if kind == 0:
            p = rng.integers(0, 4, (n, 3)).astype(float)
        elif kind == 1:
            p = rng.integers(0, 7, (n, 3)).astype(float) * 0.1 + 5.0
        elif kind == 2:
            v = rng.normal(size=(n, 3))
            p = v / np.linalg.norm(v, axis=1)[:, None]
        elif kind == 3:
            p = rng.random((n, 3))
            p[: n // 2, 2] = 0.5
        else:
            p = rng.random((n, 3))
            p[: n // 3] = p[0] + np.outer(rng.random(n // 3), p[1] - p[0])
        p = np.ascontiguousarray(p)
        t = tri.triangulate(p)
        if len(t) == 0:
            continue  # no four affinely independent points: Qhull raised, the wrapper has no cells
        _check_triangulation(hl, p, _first_occurrence_cells(p, t), n_used=len(np.unique(p, axis=0)))
        done += 1
    assert done >= 90 and tri.qhull_retries == 0


def test_triangulate_into_raw_buffers(hl):  # noqa: F811
    """the triangulators write their cells straight into a caller-owned buffer (the
    pinned staging buffer of the device loop); same cells as `triangulate`, and a buffer that is too
    small is reported as minus the capacity that is needed, with nothing written past its end."""
    from seismicmesh_b200.triangulator import get_triangulator

    rng = np.random.default_rng(11)
    for dim in (2, 3):
        tri = get_triangulator(None, dim)
        p = np.ascontiguousarray(rng.random((3000, dim)))
        ref = tri.triangulate(p)
        buf = np.full((tri.max_cells(len(p)) + 7, dim + 1), -7, dtype=np.int32)
        T = tri.triangulate_into(p, buf)
        assert T == len(ref) and np.array_equal(buf[:T], ref) and (buf[T:] == -7).all()
        small = np.full((len(ref) // 2, dim + 1), -7, dtype=np.int32)
        assert tri.triangulate_into(p, small) == -len(ref) and (small == -7).all()
        q = get_triangulator("qhull", dim)
        buf2 = np.empty((q.max_cells(len(p)), dim + 1), dtype=np.int32)
        T2 = q.triangulate_into(p, buf2)
        assert np.array_equal(_canon(buf2[:T2]), _canon(ref))


# ---- dmh_delaunay3d_mt: the same triangulation built by several host threads -----------------------

def _rows_sorted(t):
    return np.sort(np.asarray(t), axis=1)


@pytest.mark.parametrize("threads", [2, 3, 8])
def test_threads_same_cells_in_the_same_order(hl, threads):  # noqa: F811
    """Points in general position (large enough for the partitioned rounds: 8000+ rows per round): the
    threaded construction returns the cell list of the serial one, cell for cell (the column order
    inside a cell is the construction's own and may differ)."""
    from seismicmesh_b200.triangulator import BowyerWatsonTriangulator

    p = np.random.default_rng(5).random((60000, 3)) * [2.0, 1.0, 1.0]
    one = BowyerWatsonTriangulator(3, threads=1).triangulate(p)
    tri = BowyerWatsonTriangulator(3, threads=threads)
    many = tri.triangulate(p)
    assert many.dtype == np.int32 and many.flags.c_contiguous and _is_lex_sorted(many) and _column0_unbiased(many)
    assert np.array_equal(_rows_sorted(one), _rows_sorted(many))
    assert tri.qhull_retries == 0
    # raw-buffer form, and the capacity report
    buf = np.full((tri.max_cells(len(p)), 4), -7, dtype=np.int32)
    T = tri.triangulate_into(p, buf)
    assert T == len(one) and np.array_equal(_rows_sorted(buf[:T]), _rows_sorted(one)) and (buf[T:] == -7).all()
    small = np.full((1000, 4), -7, dtype=np.int32)
    assert tri.triangulate_into(p, small) == -len(one) and (small == -7).all()


def test_threads_distmesh_ball_iterates(hl):  # noqa: F811
    """What the loop feeds it: the reference's staggered lattice clipped to a ball (co-spherical
    everywhere: any valid Delaunay triangulation will do, checked exactly), then jittered and moved by
    force iterations (general position: the serial cell set)."""
    from oracle import distmesh_oracle as orc
    from seismicmesh_b200.triangulator import BowyerWatsonTriangulator

    h0 = 0.06
    spec = ("ball", {"x0": [0.0, 0.0, 0.0], "r": 1.0})
    fd = lambda x: orc.sdf(spec, x)  # noqa: E731
    fh = lambda x: np.full(len(x), h0)  # noqa: E731
    geps, deps = 0.1 * h0, np.sqrt(np.finfo(float).eps) * h0
    p0 = orc.initial_points(h0, geps, 3, np.array([[-1.0, 1.0]] * 3), fh, fd, np.empty((0, 3)))
    assert len(p0) > 16000
    one, many = BowyerWatsonTriangulator(3, threads=1), BowyerWatsonTriangulator(3, threads=4)
    t = many.triangulate(p0)
    assert np.unique(t).size == len(p0)
    _check_triangulation(hl, p0, t)
    p = p0 + np.random.default_rng(0).uniform(-0.1 * h0, 0.1 * h0, p0.shape)
    for it in range(2):
        t = many.triangulate(p)
        assert np.array_equal(_rows_sorted(t), _rows_sorted(one.triangulate(p)))
        p = orc.force_iteration(p, t, [fd], fh, h0, geps, deps)["p"]
    assert many.qhull_retries == 0


def test_threads_degenerate_inputs(hl):  # noqa: F811
    """Degenerate input through the partitioned rounds: a cubic lattice (every cube co-spherical, every
    lattice plane co-planar), duplicated rows, points on the faces of a cube."""
    from seismicmesh_b200.triangulator import BowyerWatsonTriangulator

    tri = BowyerWatsonTriangulator(3, threads=4)
    g = np.stack(np.meshgrid(np.arange(22.0), np.arange(23.0), np.arange(24.0), indexing="ij"), -1).reshape(-1, 3) * 0.1
    t = tri.triangulate(np.ascontiguousarray(g))
    assert len(t) >= 5 * 21 * 22 * 23
    _check_triangulation(hl, g, t)
    rng = np.random.default_rng(3)
    c = rng.random((12000, 3))
    for k in range(3):
        c[500 * (2 * k): 500 * (2 * k + 1), k] = 0.0
        c[500 * (2 * k + 1): 500 * (2 * k + 2), k] = 1.0
    d2 = np.ascontiguousarray(np.r_[c, c[:700]])  # 700 exact duplicates
    t = tri.triangulate(d2)
    assert _is_lex_sorted(t)
    _check_triangulation(hl, c, _first_occurrence_cells(d2, t), n_used=len(c))
    assert tri.qhull_retries == 0


def test_threads_error_paths(hl):  # noqa: F811
    T, dups, lost = C.c_int64(), C.c_int64(), C.c_int64()
    p = np.random.default_rng(1).random((20000, 3))
    out = np.empty((64, 4), np.int32)
    assert hl.dmh_delaunay3d_mt(p.ctypes.data, 20000, out.ctypes.data, 64, C.byref(T), None, None, 4) == -2
    assert T.value > 64
    flat = np.ascontiguousarray(np.c_[p[:, :2], np.zeros(len(p))])
    big = np.empty((hl.dmh_delaunay3d_max_cells(len(p)), 4), np.int32)
    assert hl.dmh_delaunay3d_mt(flat.ctypes.data, len(p), big.ctypes.data, len(big), C.byref(T), C.byref(dups), C.byref(lost), 4) == 0
    assert (T.value, lost.value) == (0, len(p))
    bad = p.copy()
    bad[77, 2] = np.inf
    assert hl.dmh_delaunay3d_mt(bad.ctypes.data, len(p), big.ctypes.data, len(big), C.byref(T), None, None, 4) == -1


def test_host_threads_split_among_ranks(monkeypatch):
    import os

    from seismicmesh_b200 import triangulator as tr

    cores = len(os.sched_getaffinity(0))
    monkeypatch.delenv("DM_HOST_THREADS", raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    want = lambda c: max(1, min(16, c - 1 if c >= 8 else c))  # noqa: E731
    assert tr.host_threads() == want(cores)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert tr.host_threads() == want(cores // 8)
    monkeypatch.setenv("DM_HOST_THREADS", "5")
    assert tr.host_threads() == 5


# ---- dmh_dt3_*: a triangulation that stays around (build, cells, insert more points, cells) ----------

@pytest.mark.parametrize("threads", [1, 4])
def test_incremental_insert_same_cells_as_rebuild(hl, threads):  # noqa: F811
    """What a slab rank does every iteration: triangulate the owned vertices, read the cells, INSERT the
    ghost vertices into the same triangulation, read the cells again.  Both cell lists must equal a
    construction from scratch (points in general position), in the canonical output order."""
    from seismicmesh_b200.triangulator import BowyerWatsonTriangulator, get_triangulator

    rng = np.random.default_rng(2)
    own = rng.random((12000, 3))
    ghosts = [rng.random((900, 3)) * [1.0, 0.12, 1.0] - [0.0, 0.12, 0.0], rng.random((700, 3)) * [1.0, 0.12, 1.0] + [0.0, 1.0, 0.0]]
    tri = BowyerWatsonTriangulator(3, threads=threads)
    dt = tri.build(own)
    t_own = dt.cells()
    assert _is_lex_sorted(t_own) and np.array_equal(_rows_sorted(t_own), _rows_sorted(tri.triangulate(own)))
    for g in ghosts:
        dt.insert(g)
    dt.insert(np.empty((0, 3)))
    t_loc = dt.cells()
    allp = np.vstack([own] + ghosts)
    assert t_loc.dtype == np.int32 and _is_lex_sorted(t_loc) and np.unique(t_loc).size == len(allp)
    assert np.array_equal(_rows_sorted(t_loc), _rows_sorted(tri.triangulate(allp)))
    dt.close()
    dt.close()  # idempotent
    if threads > 1:  # batches large enough for the partitioned passes (two ghost layers beyond flat sides of the hull)
        big = [rng.random((9000, 3)) * [1.0, 0.15, 1.0] - [0.0, 0.15, 0.0], rng.random((9000, 3)) * [1.0, 0.15, 1.0] + [0.0, 1.0, 0.0]]
        dt = tri.build(own)
        for g in big:
            dt.insert(g)
        t_big = dt.cells()
        dt.close()
        assert np.array_equal(_rows_sorted(t_big), _rows_sorted(tri.triangulate(np.vstack([own] + big))))
    # a start without four affinely independent points is redone from all points at the first insert
    flat = np.ascontiguousarray(np.c_[rng.random((40, 2)), np.zeros(40)])
    more = rng.random((300, 3))
    T, dups, lost, rc = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
    h = hl.dmh_dt3_build(flat.ctypes.data, len(flat), threads, C.byref(rc))
    assert h and rc.value == 0 and hl.dmh_dt3_points(h) == 40
    out = np.empty((hl.dmh_delaunay3d_max_cells(340), 4), np.int32)
    assert hl.dmh_dt3_cells(h, out.ctypes.data, len(out), C.byref(T), C.byref(dups), C.byref(lost)) == 0
    assert (T.value, lost.value) == (0, 40)
    assert hl.dmh_dt3_insert(h, more.ctypes.data, len(more)) == 0 and hl.dmh_dt3_points(h) == 340
    assert hl.dmh_dt3_cells(h, out.ctypes.data, len(out), C.byref(T), C.byref(dups), C.byref(lost)) == 0 and lost.value == 0
    assert np.array_equal(_rows_sorted(out[: T.value]), _rows_sorted(tri.triangulate(np.vstack((flat, more)))))
    assert hl.dmh_dt3_cells(h, out.ctypes.data, 10, C.byref(T), None, None) == -2 and T.value > 10
    bad = more.copy()
    bad[3, 0] = np.nan
    assert hl.dmh_dt3_insert(h, bad.ctypes.data, len(bad)) == -1 and hl.dmh_dt3_points(h) == 340  # nothing recorded
    hl.dmh_dt3_free(h)
    assert hl.dmh_dt3_insert(None, more.ctypes.data, 1) == -1
    # triangulators without state answer the same calls by rebuilding
    q = get_triangulator("qhull", 3).build(own[:400])
    q.insert(ghosts[0][:40])
    assert np.array_equal(_canon(q.cells()), _canon(tri.triangulate(np.vstack((own[:400], ghosts[0][:40])))))


def test_incremental_insert_random_batches_and_duplicates(hl):  # noqa: F811
    """Several batches of random size into one triangulation, some rows exact duplicates of rows already in
    (a ghost vertex that coincides with an owned one), some on the hull's flat sides: after every batch the
    cells are a valid Delaunay triangulation of the distinct points (exact checks) and, the points being in
    general position apart from the duplicates, the cell SET of a construction from scratch."""
    from seismicmesh_b200.triangulator import BowyerWatsonTriangulator

    rng = np.random.default_rng(17)
    tri = BowyerWatsonTriangulator(3, threads=3)
    for trial in range(4):
        pts = [rng.random((int(rng.integers(5, 400)), 3))]
        dt = tri.build(pts[0])
        for b in range(int(rng.integers(1, 5))):
            q = rng.random((int(rng.integers(1, 300)), 3)) * [1.0, 0.3, 1.0] + [0.0, rng.choice([-0.3, 0.35, 1.0]), 0.0]
            allp = np.vstack(pts)
            k = min(len(q) // 4, len(allp))
            if k:
                q[:k] = allp[rng.choice(len(allp), k, replace=False)]  # exact duplicates of rows already in
            dt.insert(q)
            pts.append(np.ascontiguousarray(q))
            allp = np.vstack(pts)
            t = dt.cells()
            assert _is_lex_sorted(t)
            uniq = np.unique(allp, axis=0)
            first = _first_occurrence_cells(allp, t)
            _, fi = np.unique(allp, axis=0, return_index=True)
            remap = -np.ones(len(allp), dtype=np.int64)
            remap[np.sort(fi)] = np.arange(len(fi))
            dense = remap[first]
            assert (dense >= 0).all()
            pd_ = allp[np.sort(fi)]
            _check_triangulation(hl, pd_, _canon(dense).astype(np.int32), n_used=len(uniq))
            ref = _first_occurrence_cells(allp, tri.triangulate(allp))
            assert np.array_equal(_canon(first), _canon(ref))
        dt.close()
    assert tri.qhull_retries == 0
