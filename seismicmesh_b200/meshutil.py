"""Host-side mesh clean-up that runs ONCE at termination (out of the hot loop; SURVEY section 8
row "next #2").  Own NumPy/SciPy implementations of the behaviour of the reference's
geometry/utils.py helpers that `_termination` needs (mesh_generator.py:655-677):
fix_mesh (:204-249), simp_vol (:175-199), simp_qual (:252-274), boundary queries (:310-437),
delete_boundary_entities (:440-468) and laplacian2_fixed_point (:494-547); plus the optional
`perform_checks` pass: linter (:820-872) over do_any_overlap (:745-817) and is_manifold (:634-654).
"""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu

__all__ = [
    "simp_vol", "simp_qual", "fix_mesh", "get_edges", "get_facets", "get_boundary_edges",
    "get_boundary_facets", "get_boundary_vertices", "get_boundary_entities",
    "delete_boundary_entities", "laplacian2_fixed_point", "laplacian2", "calc_re_ratios", "get_winded_boundary_edges",
    "vertex_in_entity3", "get_centroids", "vertex_to_entities",
    "do_any_overlap", "is_manifold", "linter", "unique_rows",
]


def simp_vol(p, t):
    """Signed simplex volumes (area in 2-D)."""
    dim = p.shape[1]
    a = p[t[:, 1]] - p[t[:, 0]]
    if dim == 1:
        return a
    b = p[t[:, 2]] - p[t[:, 0]]
    if dim == 2:
        return (a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]) / 2
    if dim == 3:
        c = p[t[:, 3]] - p[t[:, 0]]
        return np.einsum("ij,ij->i", np.cross(a, b), c) / 6
    raise NotImplementedError


def simp_qual(p, t):
    """2 * inradius / circumradius of the triangle spanned by the first three vertices."""
    assert p.ndim == 2 and t.ndim == 2 and p.shape[1] + 1 == t.shape[1]

    def norm(v):
        return np.sqrt((v**2).sum(1))

    a = norm(p[t[:, 1]] - p[t[:, 0]])
    b = norm(p[t[:, 2]] - p[t[:, 0]])
    c = norm(p[t[:, 2]] - p[t[:, 1]])
    s = a + b + c
    prod = (b + c - a) * (c + a - b) * (a + b - c)
    r = 0.5 * np.sqrt(prod / s)
    R = a * b * c / np.sqrt(s * prod)
    return 2 * r / R


def _unique_rows(a, return_index=False, return_inverse=False, return_counts=False, rows_sorted=False):
    """`np.unique(a, axis=0, ...)` (rows in lexicographic order, index of the first occurrence,
    inverse, counts) without NumPy's structured-dtype sort, which dominates the termination step:
    integer rows are packed into one int64 key per row when they fit, everything else goes through
    a stable lexsort."""
    a = np.ascontiguousarray(a)
    n, m = a.shape
    if n == 0:
        return np.unique(a, axis=0, return_index=return_index, return_inverse=return_inverse, return_counts=return_counts)
    if rows_sorted and not (return_index or return_inverse) and a.dtype.kind in "iu" and m in (2, 3, 4) and n < 2**31:
        # the termination path's case (cells / facets / edges with ascending ids per row): native code
        lo, hi = int(a.min()), int(a.max())
        if lo >= 0 and hi < min(2**31 - 1, 64 * n + 4096):  # (ids about as dense as vertex ids are: the code buckets by id)
            import ctypes as C

            from ._hostlib import lib as _host
            from .triangulator import host_threads

            rows = np.array(a, dtype=np.int32, order="C")  # (a copy: collapsed in place)
            counts = np.empty(n, dtype=np.int32) if return_counts else None
            nu = C.c_int64(0)
            rc = _host().dmh_sort_unique_rows_i32(rows.ctypes.data, n, m, hi + 1, counts.ctypes.data if return_counts else None,
                                                  C.byref(nu), host_threads())
            if rc != 0:
                raise RuntimeError(f"dmh_sort_unique_rows_i32 failed with code {rc}")
            u = rows[: nu.value].astype(a.dtype, copy=False)
            return (u, counts[: nu.value].astype(np.intp)) if return_counts else u
    bits = 63 // m
    if a.dtype.kind in "iu" and a.min() >= 0 and int(a.max()) < (1 << bits):
        key = a[:, 0].astype(np.int64)
        for j in range(1, m):
            key = (key << bits) | a[:, j].astype(np.int64)
        order = np.argsort(key, kind="stable")
        ks = key[order]
        first = np.concatenate(([True], ks[1:] != ks[:-1]))
    else:
        order = np.lexsort(a.T[::-1])  # stable; last key is the primary one
        srt = a[order]
        first = np.concatenate(([True], (srt[1:] != srt[:-1]).any(axis=1)))
    idx = order[first]
    out = [a[idx]]
    if return_index:
        out.append(idx)
    if return_inverse:
        inv = np.empty(n, dtype=np.intp)
        inv[order] = np.cumsum(first) - 1
        out.append(inv)
    if return_counts:
        pos = np.flatnonzero(first)
        out.append(np.diff(np.concatenate((pos, [n]))))
    return out[0] if len(out) == 1 else tuple(out)


def fix_mesh(p, t, ptol=2e-13, dim=2, delete_unused=False, fix_orientation=True):
    """Merge duplicate vertices / cells, optionally drop unused vertices, make cells CCW."""
    snap = (p.max(0) - p.min(0)).max() * ptol
    _, ix, jx = _unique_rows(np.round(p / snap) * snap, return_index=True, return_inverse=True)
    jx = np.asarray(jx).ravel()
    p = p[ix]
    t = jx[t]
    t = _unique_rows(np.sort(t, axis=1), rows_sorted=True)
    if delete_unused:
        used, jx = np.unique(t, return_inverse=True)
        t = np.asarray(jx).reshape(t.shape)
        p = p[used, :]
    if fix_orientation:
        flip = simp_vol(p, t) < 0
        t[flip, :2] = t[flip, 1::-1]
    return p, t, jx


def get_edges(t, dim=2):
    t = np.asarray(t)
    pairs = [[0, 1], [0, 2], [1, 2]] if dim == 2 else [[0, 1], [1, 2], [2, 0], [0, 3], [1, 3], [2, 3]]
    return t[:, pairs].reshape((-1, 2))


def get_facets(t):
    return np.asarray(t)[:, [[0, 1, 3], [1, 2, 3], [2, 0, 3], [1, 2, 0]]].reshape((-1, 3))


def _once(rows, count):
    u, c = _unique_rows(np.sort(rows, axis=1), return_counts=True, rows_sorted=True)
    return u[c == count]


def get_boundary_edges(t, dim=2):
    return _once(get_edges(t, dim=dim), dim - 1)


def get_boundary_facets(t):
    if np.asarray(t).shape[1] < 4:
        raise ValueError("Only works for triangles")
    return _once(get_facets(t), 1)


def get_boundary_vertices(t, dim=2):
    if dim == 2:
        b = get_boundary_edges(t)
    elif dim == 3:
        b = get_boundary_facets(t)
    else:
        raise ValueError("Dimension not supported.")
    return np.unique(b.reshape(-1))


def get_boundary_entities(p, t, dim=2):
    """Indices of cells incident to at least one boundary vertex."""
    bv = get_boundary_vertices(t, dim=dim)
    mark = np.zeros(len(p), dtype=bool)
    mark[bv] = True
    return np.nonzero(mark[t].any(axis=1))[0]


def delete_boundary_entities(p, t, dim=2, min_qual=0.10, verbose=1):
    qual = simp_qual(p, t)
    bele = get_boundary_entities(p, t, dim=dim)
    bad = qual[bele] < min_qual
    if verbose:
        print("Deleting " + str(np.sum(bad)) + " poor quality boundary entities...", flush=True)
    t = np.delete(t, bele[bad], axis=0)
    p, t, _ = fix_mesh(p, t, delete_unused=True, dim=dim)
    return p, t


def laplacian2_fixed_point(p, t):
    """Laplacian smoothing as ONE linear solve with Dirichlet boundary vertices: every interior
    vertex goes to the (edge-multiplicity weighted) average of its neighbours."""
    if p.ndim != 2:
        raise NotImplementedError("Laplacian smoothing only works in 2D for now")
    n = len(p)
    i0 = np.concatenate([t[:, 1], t[:, 2], t[:, 0]])
    i1 = np.concatenate([t[:, 2], t[:, 0], t[:, 1]])
    ones = np.ones(len(i0))
    rows = np.concatenate([i0, i1, i0, i1])
    cols = np.concatenate([i0, i1, i1, i0])
    vals = np.concatenate([ones, ones, -ones, -ones])
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    bnd = get_boundary_vertices(t)
    interior = np.ones(n, dtype=bool)
    interior[bnd] = False
    # Dirichlet rows: identity
    D = sp.diags(interior.astype(float))
    A = (D @ A + sp.diags((~interior).astype(float))).tocsc()
    rhs = np.zeros((n, 2))
    rhs[bnd] = p[bnd]
    lu = splu(A)
    return np.column_stack([lu.solve(rhs[:, 0]), lu.solve(rhs[:, 1])]), t


def unique_rows(a, return_index=False, return_inverse=False):
    """`geometry.unique_rows` (geometry/utils.py:141-172): unique rows in lexicographic order."""
    out = _unique_rows(np.asarray(a), return_index=return_index, return_inverse=return_inverse)
    return out


def get_centroids(p, t, dim=2):
    """p[t].sum(1)/(dim+1) (geometry/utils.py:729-742)."""
    return p[t].sum(1) / (dim + 1)


def vertex_to_entities(p, t, dim=2):
    """CSR map vertex -> incident entities: (vtoe, ptr) as geometry/utils.py:577-602 returns it."""
    t = np.asarray(t)
    n = len(p)
    flat = t.ravel()
    order = np.argsort(flat, kind="stable")
    vtoe = (order // t.shape[1]).astype(np.int64)
    ptr = np.concatenate([[0], np.cumsum(np.bincount(flat, minlength=n))]).astype(np.int64)
    return vtoe, ptr


def do_any_overlap(p, t, dim=2):
    """Pairs (ie, ele) where the centroid of entity `ie` lies inside an entity `ele` of its 1-ring
    (geometry/utils.py:745-817; the point-in-entity tests are vertex_in_entity2 :657-679, closed
    barycentric bounds, and vertex_in_entity3 :682-726, five determinants of one strict sign).
    Vectorised over all (entity, ring neighbour) pairs instead of the reference's Python loops."""
    t = np.asarray(t)
    T, c = t.shape
    if T == 0:
        return []
    vtoe, ptr = vertex_to_entities(p, t, dim=dim)
    cnt = (ptr[1:] - ptr[:-1])[t]                      # (T, c) ring sizes per corner
    tot = cnt.sum(axis=1)
    ie = np.repeat(np.arange(T), tot)
    # neighbour ids: for entity ie, concatenate vtoe[ptr[v]:ptr[v+1]] over its vertices (reference order)
    starts = ptr[:-1][t].ravel()
    lens = cnt.ravel()
    seg = np.repeat(np.arange(len(lens)), lens)
    within = np.arange(lens.sum()) - np.repeat(np.cumsum(lens) - lens, lens)
    ele = vtoe[starts[seg] + within]
    m = ie != ele
    ie, ele = ie[m], ele[m]
    cen = get_centroids(p, t, dim=dim)[ie]
    q = p[t[ele]]                                      # (M, c, dim)
    if dim == 2:
        x, y = cen[:, 0], cen[:, 1]
        x1, y1, x2, y2, x3, y3 = q[:, 0, 0], q[:, 0, 1], q[:, 1, 0], q[:, 1, 1], q[:, 2, 0], q[:, 2, 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            den = (y2 - y3) * (x1 - x3) + (x3 - x2) * (y1 - y3)
            a = ((y2 - y3) * (x - x3) + (x3 - x2) * (y - y3)) / den
            b = ((y3 - y1) * (x - x3) + (x1 - x3) * (y - y3)) / den
        cc = 1 - a - b
        inside = (0 <= a) & (a <= 1) & (0 <= b) & (b <= 1) & (0 <= cc) & (cc <= 1)
    else:
        def det4(rows):
            A = np.concatenate([rows, np.ones(rows.shape[:2] + (1,))], axis=2)
            return np.sign(np.linalg.det(A))

        s0 = det4(q)
        inside = s0 != 0
        for k in range(4):
            r = q.copy()
            r[:, k, :] = cen
            inside &= det4(r) == s0
    return [(int(a_), int(b_)) for a_, b_ in zip(ie[inside], ele[inside])]


def is_manifold(p, t, dim=2):
    """geometry/utils.py:634-654: #boundary-edge endpoints == 2 * #boundary vertices."""
    bedges = get_boundary_edges(t, dim=dim)
    if bedges.size != p[np.unique(bedges), :].size:
        print("Mesh has a non-manifold boundary...", flush=True)
        return False
    return True


def linter(p, t, dim=2, min_qual=0.10):
    """`perform_checks=True` clean-up (geometry/utils.py:820-872): of every overlapping pair the
    lower-quality entity goes, then fix_mesh, poor boundary entities (2-D) and the manifold check."""
    print("Performing mesh linting...", flush=True)
    qual = simp_qual(p, t)
    pairs = do_any_overlap(p, t, dim=dim)
    delete = np.unique(np.array([pr[int(np.argmin(qual[list(pr)]))] for pr in pairs], dtype=int))
    print("Deleting " + str(len(delete)) + " overlapped entities", flush=True)
    t = np.delete(t, delete, axis=0)
    p, t, _ = fix_mesh(p, t, delete_unused=True, dim=dim)
    if dim == 2:
        p, t = delete_boundary_entities(p, t, min_qual=min_qual)
        is_manifold(p, t)
    qual = simp_qual(p, t)
    print("There are " + str(len(p)) + " vertices and " + str(len(t)) + " elements in the mesh", flush=True)
    print("The minimum element quality is " + str(np.amin(qual)), flush=True)
    return p, t


# ----------------------------------------------------------------------------------------------
# Small helpers of the reference's geometry namespace that its own tests call (tests/test_geometry.py,
# test_geometry2.py, test_ptin.py): mesh-quality ratio, boundary loop, iterative smoothing, point-in-tet.
# Host NumPy; not on the device path.
# ----------------------------------------------------------------------------------------------
def calc_re_ratios(vertices, entities, dim=2):
    """Circumradius over shortest edge of every simplex (geometry/utils.py:17-54; the reference takes the
    circumballs from CGAL, here they come from the perpendicular-bisector system)."""
    p = np.asarray(vertices, dtype=np.float64)
    t = np.asarray(entities)
    if dim not in (2, 3):
        raise ValueError("Dimension invalid")
    a = p[t[:, 0]]
    rows = p[t[:, 1:]] - a[:, None, :]                    # (T, dim, dim): edges from vertex 0
    rhs = 0.5 * np.einsum("tij,tij->ti", rows, rows)      # |e_i|^2 / 2
    centre = np.linalg.solve(rows, rhs[..., None])[..., 0]  # circumcentre relative to vertex 0
    radius = np.sqrt(np.einsum("ti,ti->t", centre, centre))
    pairs = [(0, 1), (1, 2), (2, 0)] if dim == 2 else [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]
    lengths = np.stack([np.sqrt(((p[t[:, i]] - p[t[:, j]]) ** 2).sum(1)) for i, j in pairs])
    return radius / lengths.min(axis=0)


def get_winded_boundary_edges(entities):
    """The boundary edges in the order a walk along the boundary meets them, starting from the first one
    (geometry/utils.py:329-361): from the current vertex take the unvisited boundary edge with the smallest
    row index that contains it; rows keep their stored orientation."""
    be = get_boundary_edges(entities)
    if len(be) == 0:
        return be
    incident = {}
    for r, (u, v) in enumerate(be):
        incident.setdefault(int(u), []).append(r)
        incident.setdefault(int(v), []).append(r)
    visited = np.zeros(len(be), dtype=bool)
    order = [0]
    visited[0] = True
    cur = int(be[0, 1])
    while True:
        nxt = [r for r in incident.get(cur, []) if not visited[r]]
        if not nxt:
            break
        r = min(nxt)
        visited[r] = True
        order.append(r)
        u, v = int(be[r, 0]), int(be[r, 1])
        cur = v if u == cur else u
    return be[order]


def laplacian2(vertices, entities, max_iter=20, tol=0.01, verbose=1, pfix=None):
    """Iterated Laplacian smoothing (geometry/utils.py:550-631): every non-boundary vertex moves to the average
    of its neighbours (an edge shared by two triangles counts twice) until the largest relative change of an
    edge length drops below `tol`; `pfix` = coordinates whose nearest vertices stay put as well."""
    p = np.asarray(vertices, dtype=np.float64)
    t = np.asarray(entities)
    if p.ndim != 2 or p.shape[1] != 2:
        raise NotImplementedError("Laplacian smoothing only works in 2D for now")
    n = len(p)
    i = t[:, [0, 0, 1, 1, 2, 2]].ravel()
    j = t[:, [1, 2, 0, 2, 0, 1]].ravel()
    S = sp.coo_matrix((np.ones(len(i)), (i, j)), shape=(n, n)).tocsr()
    W = np.asarray(S.sum(1)).ravel()
    if np.any(W == 0):
        print("Invalid mesh. Disjoint vertices found. Returning", flush=True)
        print(np.argwhere(W == 0), flush=True)
        return vertices, entities
    fixed = get_boundary_vertices(t)
    if pfix is not None:
        near = [int(np.argmin(((p - np.asarray(f)) ** 2).sum(1))) for f in np.atleast_2d(pfix)]
        fixed = np.concatenate((fixed, np.asarray(near, dtype=fixed.dtype)))
    edge = get_edges(t)
    eps = np.finfo(float).eps

    def lengths(q):
        return np.maximum(np.sqrt(((q[edge[:, 0]] - q[edge[:, 1]]) ** 2).sum(1)), eps)

    L = lengths(p)
    for it in range(max_iter):
        q = (S @ p) / W[:, None]
        q[fixed] = p[fixed]
        p = q
        Ln = lengths(p)
        if np.amax((Ln - L) / Ln) < tol:
            if verbose:
                print("Movement tolerance reached after " + str(it) + " iterations..exiting", flush=True)
            break
        L = Ln
    return p, entities


def vertex_in_entity3(vertex, entity):
    """Is the point inside the tetrahedron given as 12 coordinates (geometry/utils.py:682-726)?  True when
    replacing each corner in turn by the point never flips the orientation."""
    q = np.asarray(vertex, dtype=np.float64)
    c = np.asarray(entity, dtype=np.float64).reshape(4, 3)

    def orient(m):
        return np.sign(np.linalg.det(np.hstack([m, np.ones((4, 1))])))

    s0 = orient(c)
    if s0 == 0:
        return False
    for k in range(4):
        m = c.copy()
        m[k] = q
        if orient(m) != s0:
            return False
    return True
