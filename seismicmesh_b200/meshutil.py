"""Host-side mesh clean-up that runs ONCE at termination (out of the hot loop; SURVEY section 8
row "next #2").  Own NumPy/SciPy implementations of the behaviour of the reference's
geometry/utils.py helpers that `_termination` needs (mesh_generator.py:655-677):
fix_mesh (:204-249), simp_vol (:175-199), simp_qual (:252-274), boundary queries (:310-437),
delete_boundary_entities (:440-468) and laplacian2_fixed_point (:494-547).
"""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu

__all__ = [
    "simp_vol", "simp_qual", "fix_mesh", "get_edges", "get_facets", "get_boundary_edges",
    "get_boundary_facets", "get_boundary_vertices", "get_boundary_entities",
    "delete_boundary_entities", "laplacian2_fixed_point",
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


def _unique_rows(a, return_index=False, return_inverse=False, return_counts=False):
    """`np.unique(a, axis=0, ...)` (rows in lexicographic order, index of the first occurrence,
    inverse, counts) without NumPy's structured-dtype sort, which dominates the termination step:
    integer rows are packed into one int64 key per row when they fit, everything else goes through
    a stable lexsort."""
    a = np.ascontiguousarray(a)
    n, m = a.shape
    if n == 0:
        return np.unique(a, axis=0, return_index=return_index, return_inverse=return_inverse, return_counts=return_counts)
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
    t = _unique_rows(np.sort(t, axis=1))
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
    u, c = _unique_rows(np.sort(rows, axis=1), return_counts=True)
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
