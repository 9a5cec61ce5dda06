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
