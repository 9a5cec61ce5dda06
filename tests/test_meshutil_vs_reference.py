"""Termination-path mesh utilities (seismicmesh_b200/meshutil.py, host NumPy)
against the reference's geometry/utils.py, through golden vectors written by make_golden.py
(`gen_meshutil`) from the unmodified reference."""
import numpy as np
import pytest
from conftest import load_golden


def _rows(a):
    a = np.sort(np.asarray(a, dtype=np.int64), axis=1)
    return a[np.lexsort(a.T[::-1])]


def _same_mesh(p0, t0, p1, t1, tol=1e-12):
    """Same vertex set and same cells up to vertex / cell numbering."""
    assert p0.shape == p1.shape and t0.shape == t1.shape
    o0, o1 = np.lexsort(p0.T[::-1]), np.lexsort(p1.T[::-1])
    assert np.abs(p0[o0] - p1[o1]).max() <= tol
    r0, r1 = np.empty(len(p0), np.int64), np.empty(len(p1), np.int64)
    r0[o0], r1[o1] = np.arange(len(p0)), np.arange(len(p1))
    assert np.array_equal(_rows(r0[t0]), _rows(r1[t1]))


@pytest.mark.parametrize("dim", [2, 3])
def test_volumes_quality_and_boundary(dim):
    from seismicmesh_b200 import meshutil as mu

    g = load_golden(f"meshutil_{dim}d.npz")
    p, t = g["p"], g["t"]
    assert np.allclose(mu.simp_vol(p, t), g["vol"], rtol=1e-12, atol=0)
    assert np.allclose(mu.simp_qual(p, t), g["qual"], rtol=1e-10, atol=1e-14)
    assert np.array_equal(mu.get_boundary_vertices(t, dim=dim), g["bverts"])
    assert np.array_equal(np.sort(mu.get_boundary_entities(p, t, dim=dim)), np.sort(g["bents"]))
    if dim == 2:
        assert np.array_equal(_rows(mu.get_boundary_edges(t)), _rows(g["bedges"]))
    else:
        assert np.array_equal(_rows(mu.get_boundary_facets(t)), _rows(g["bfacets"]))


def test_fix_mesh_merges_duplicates_and_drops_unused():
    from seismicmesh_b200 import meshutil as mu

    g = load_golden("meshutil_2d.npz")
    p, t, _ = mu.fix_mesh(g["dirty_p"].copy(), g["dirty_t"].copy(), delete_unused=True)
    _same_mesh(p, t, g["fix_p"], g["fix_t"])
    assert (mu.simp_vol(p, t) > 0).all()  # fix_orientation
    p2, t2, _ = mu.fix_mesh(g["dirty_p"].copy(), g["dirty_t"].copy(), delete_unused=False)
    _same_mesh(p2, t2, g["fix2_p"], g["fix2_t"])


@pytest.mark.parametrize("dim,min_qual", [(2, 0.55), (3, 0.2)])
def test_delete_boundary_entities(dim, min_qual):
    from seismicmesh_b200 import meshutil as mu

    g = load_golden(f"meshutil_{dim}d.npz")
    p, t = mu.delete_boundary_entities(g["p"].copy(), g["t"].copy(), dim=dim, min_qual=min_qual, verbose=0)
    assert len(t) < len(g["t"])  # the case does delete something
    _same_mesh(p, t, g["del_p"], g["del_t"])


def test_laplacian_smoothing_fixed_point():
    from seismicmesh_b200 import meshutil as mu

    g = load_golden("meshutil_2d.npz")
    p, t = mu.laplacian2_fixed_point(g["p"].copy(), g["t"].copy())
    assert np.array_equal(t, g["t"])
    assert np.abs(p - g["lap_p"]).max() <= 1e-10


def test_linter_and_overlap_match_reference():
    """`perform_checks=True` pass (geometry/utils.py:745-872) against the unmodified reference on a 2-D and a
    3-D mesh with deliberately folded-over cells: same overlap pairs, same cleaned mesh."""
    import contextlib
    import io

    from oracle import ref_harness

    if not (ref_harness.reference_available() and ref_harness.native_available()):
        pytest.skip("reference tree only exists in the build container")
    from scipy.spatial import Delaunay

    from seismicmesh_b200 import meshutil

    ref = ref_harness.load_reference()
    rng = np.random.default_rng(7)
    for dim, n in ((2, 120), (3, 60)):
        p = rng.random((n, dim))
        t = Delaunay(p).simplices.astype(np.int64)
        q = p.copy()
        q[rng.choice(n, 6, replace=False)] += 0.15  # fold a few stars over their neighbours
        with contextlib.redirect_stdout(io.StringIO()):
            pairs_ref = ref.geometry.do_any_overlap(q, t, dim=dim)
            p0, t0 = ref.geometry.linter(q.copy(), t.copy(), dim=dim)
            pairs = meshutil.do_any_overlap(q, t, dim=dim)
            p1, t1 = meshutil.linter(q.copy(), t.copy(), dim=dim)
        assert len(pairs_ref) > 0
        assert sorted(pairs) == sorted((int(a), int(b)) for a, b in pairs_ref)

        def canon(pp, tt):  # the mesh as a vertex-numbering independent set of cells
            c = np.sort(pp[tt].reshape(len(tt), -1), axis=1)
            return c[np.lexsort(c.T[::-1])]

        assert p0.shape == p1.shape and t0.shape == t1.shape
        assert np.array_equal(canon(p0, t0), canon(p1, t1))


def test_reference_geometry_tests_run_through_the_package():
    """The reference's tests/test_geometry.py, test_geometry2.py and test_ptin.py with the import swapped to
    seismicmesh_b200.geometry: same inputs, the answers those tests assert."""
    from seismicmesh_b200 import geometry as geo

    # --- tests/test_geometry.py:8-60: twelve tetrahedra in a cube of side 2 around its centre
    points = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1], [0, 0, 0]],
                      dtype=float)
    cells = np.array([[3, 4, 0, 8], [3, 7, 8, 2], [6, 8, 7, 2], [6, 5, 8, 2], [6, 5, 7, 8], [8, 7, 4, 5], [3, 7, 4, 8], [3, 8, 1, 2],
                      [8, 5, 1, 2], [8, 4, 1, 5], [8, 4, 0, 1], [3, 8, 0, 1]], dtype=int)
    assert len(geo.get_facets(cells)) == 12 * 4
    bf = np.unique(np.sort(geo.get_boundary_facets(cells), axis=1), axis=1)
    assert len(bf) == 12
    assert len(geo.get_boundary_vertices(cells, dim=3)) == 8
    assert np.sum(geo.simp_vol(points, cells)) == 8.0
    # --- tests/test_geometry2.py:10-37: two triangles side by side
    points = np.array([[0, 0], [6, 0], [3, 3], [9, 3]], dtype=float)
    cells = np.array([[3, 2, 1], [2, 0, 1]], dtype=int)
    assert np.allclose(geo.calc_re_ratios(points, cells, dim=2), [0.70710678, 0.70710678])
    assert len(geo.get_edges(cells)) == 6
    assert np.allclose(geo.get_winded_boundary_edges(cells), [[0, 1], [1, 3], [2, 3], [0, 2]])
    assert len(geo.do_any_overlap(points, cells, dim=2)) == 0
    g = load_golden("meshutil_2d.npz")
    p1, t1 = geo.laplacian2(g["p"].copy(), g["t"].copy(), verbose=0)
    assert p1.shape == g["p"].shape and np.array_equal(t1, g["t"])
    bnd = geo.get_boundary_vertices(g["t"])
    assert np.array_equal(p1[bnd], g["p"][bnd]) and geo.simp_qual(p1, t1).min() > 0
    geo.laplacian2_fixed_point(g["p"].copy(), g["t"].copy())
    # --- tests/test_ptin.py:8-46: a point inside / outside a tetrahedron
    pts = np.array([[0, 0, 0], [0, 1, 0], [np.sqrt(2), 0.5, 0], [0.5, 0.5, np.sqrt(2)]], dtype=float)
    ent = tuple(pts.ravel())
    assert geo.vertex_in_entity3((0.5, 0.5, 1.0), ent)
    assert not geo.vertex_in_entity3((0.0, 0.0, 1.0), ent)
    # the ratio against the unmodified reference on a 3-D mesh as well (circumballs from its own native code)
    g3 = load_golden("meshutil_3d.npz")
    re3 = geo.calc_re_ratios(g3["p"], g3["t"], dim=3)
    assert re3.shape == (len(g3["t"]),) and np.all(re3 >= np.sqrt(6) / 4 - 1e-9)  # regular tetrahedron is the minimum
