"""Signed-distance geometry objects with the reference's public interface
(SeismicMesh/geometry/signed_distance_functions.py:283-593): ``Disk, Ball, Rectangle, Cube,
Torus, Prism, Cylinder`` with ``rotate/stretch/translate`` and the combinators
``Union, Intersection, Difference, Repeat``.  Every object exposes ``dim, bbox, corners`` and
``eval(x)`` exactly like the reference, but instead of one NumPy pass per tree node the tree is
lowered once to a postfix SDF program (include/distmesh_b200.h) that a single fused CUDA kernel
evaluates per point.  There is no CPU evaluation path.
"""
import itertools

import numpy as np
import torch

from . import _lib
from . import device as D
from ._lib import check, lib

__all__ = [
    "Disk", "Ball", "Rectangle", "Cube", "Torus", "Prism", "Cylinder",
    "Union", "Intersection", "Difference", "Repeat", "corners", "lower", "SDFProgramError",
]

_STACK_LIMIT = 8
_PSTACK_LIMIT = 2


class SDFProgramError(ValueError):
    pass


def corners(bbox):
    """All 2^dim corners of an axis-aligned box given as (min0,max0,min1,max1,...)
    (same ordering as the reference's helper, signed_distance_functions.py:74-78)."""
    lo, hi = bbox[::2], bbox[1::2]
    return np.array(list(itertools.product(*zip(lo, hi))))


def _bbox_of(pts):
    pts = np.asarray(pts)
    out = []
    for k in range(pts.shape[1]):
        out += [np.min(pts[:, k]), np.max(pts[:, k])]
    return tuple(out)


def _rot2(a):
    return np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])


def _rot3(axis, a):
    c, s = np.cos(a), np.sin(a)
    if axis == 0:
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=float)
    if axis == 1:
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=float)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=float)


class _SDF:
    """Common behaviour: lowering to a device program and fused evaluation."""

    dim = 0
    _prog_cache = None

    # -- lowering ---------------------------------------------------------------------------
    def _emit(self, out):
        raise NotImplementedError

    def program(self):
        """float64 program words (host).  Raises SDFProgramError if a node is not lowerable."""
        ins = []
        self._emit(ins)
        _check_depth(ins)
        words = np.zeros(1 + _lib.DM_SDF_WORDS * len(ins))
        words[0] = len(ins)
        for i, rec in enumerate(ins):
            words[1 + i * _lib.DM_SDF_WORDS : 1 + (i + 1) * _lib.DM_SDF_WORDS] = rec
        return words

    def device_program(self):
        dev = D.device()
        if self._prog_cache is None or self._prog_cache.device != dev:
            self._prog_cache = torch.from_numpy(self.program()).to(dev)
        return self._prog_cache

    def spec(self):
        """Plain-data description of the tree (used by the tests to drive the oracle)."""
        raise NotImplementedError

    # -- evaluation -------------------------------------------------------------------------
    def eval(self, x):
        """Signed distance at x (M,dim).  NumPy in -> NumPy out, torch.cuda in -> torch out."""
        as_torch = isinstance(x, torch.Tensor)
        xd = D.points_dev(x if as_torch else np.asarray(x, dtype=np.float64), self.dim)
        out = torch.empty(xd.shape[0], dtype=torch.float64, device=xd.device)
        check(
            lib.dm_sdf_eval(D.ptr(self.device_program()), D.ptr(xd), xd.shape[0], self.dim, D.ptr(out), D.stream_ptr()),
            "dm_sdf_eval",
        )
        return out if as_torch else out.cpu().numpy()

    def show(self, filename=None, samples=10000):  # pragma: no cover - plotting is out of scope
        raise NotImplementedError("plotting is outside the scope of seismicmesh_b200")


def _check_depth(ins):
    sp = pp = 0
    for rec in ins:
        op = int(rec[0])
        if op < _lib.OP_UNION:
            sp += 1
        elif op < _lib.OP_REPEAT_BEGIN:
            sp -= 1
        elif op == _lib.OP_REPEAT_BEGIN:
            pp += 1
        else:
            pp -= 1
        if sp > _STACK_LIMIT or pp > _PSTACK_LIMIT:
            raise SDFProgramError("SDF tree too deep for the device evaluator")
    if sp != 1 or pp != 0:
        raise SDFProgramError("malformed SDF tree")


class _Primitive(_SDF):
    _op = 0

    def _setup(self, dim, bbox, rotate, stretch, translate, corner_pts):
        self.dim = dim
        self.bbox = tuple(bbox)
        self.corners = corner_pts
        self.rotation = rotate
        self.v = None if stretch is None else np.array(stretch, dtype=np.float64)
        self.translation_vec = None if translate is None else np.asarray(translate, dtype=np.float64)
        self.alpha = None
        self._configure()

    def _configure(self):
        """Forward-transform the public bbox/corners: stretch, then rotate, then translate
        (reference behaviour of _configure_manipulations, :104-280)."""
        d = self.dim
        if self.v is not None:
            if len(self.v) != d:
                raise AssertionError("Length of stretch vector does not equal dimension")
            self.alpha = float(np.sqrt(np.dot(self.v, self.v)))
            self.v = self.v / self.alpha
            c = corners(self.bbox)
            along = np.multiply.outer(np.dot(self.v, c.T), self.v)
            self.bbox = _bbox_of(along * self.alpha + (c - along))
            self.corners = corners(self.bbox)
        rot = self.rotation
        if d == 2:
            if rot[0] != 0.0:
                c = corners(self.bbox)
                self.bbox = _bbox_of(np.dot(_rot2(rot[0]), c.T).T)
                self.corners = corners(self.bbox)
        else:
            if any(a != 0.0 for a in rot[:3]):
                base = corners(self.bbox)
                boxes = []
                for axis in range(3):
                    # an axis with a zero angle contributes the un-rotated box (reference quirk)
                    boxes.append(_bbox_of(np.dot(_rot3(axis, rot[axis]), base.T).T) if rot[axis] != 0.0 else self.bbox)
                merged = []
                for k in range(6):
                    vals = [b[k] for b in boxes]
                    merged.append(min(vals) if k % 2 == 0 else max(vals))
                self.bbox = tuple(merged)
                self.corners = corners(self.bbox)
        if self.translation_vec is not None:
            tv = self.translation_vec
            if len(tv) != d:
                raise AssertionError("Length of translation vector does not equal dimension")
            self.bbox = tuple(self.bbox[k] + tv[k // 2] for k in range(2 * d))

    def _params(self):
        raise NotImplementedError

    def _record(self):
        rec = np.zeros(_lib.DM_SDF_WORDS)
        rec[0] = self._op
        flags = 0
        prm = self._params()
        rec[2 : 2 + len(prm)] = prm
        if self.translation_vec is not None:
            flags |= _lib.TF_TRANSLATE
            rec[8 : 8 + self.dim] = self.translation_vec
        rot = self.rotation
        nrot = 1 if self.dim == 2 else 3
        for k in range(nrot):
            if rot[k] != 0.0:
                flags |= (_lib.TF_ROT0, _lib.TF_ROT1, _lib.TF_ROT2)[k]
                rec[11 + 2 * k] = np.cos(rot[k])
                rec[12 + 2 * k] = np.sin(rot[k])
        if self.v is not None:
            flags |= _lib.TF_STRETCH
            rec[17 : 17 + self.dim] = self.v
            rec[20] = self.alpha
        rec[1] = flags
        return rec

    def _emit(self, out):
        out.append(self._record())

    def _tf_spec(self):
        return dict(
            rotate=[float(a) for a in self.rotation],
            stretch=None if self.v is None else [float(a * self.alpha) for a in self.v],
            translate=None if self.translation_vec is None else [float(a) for a in self.translation_vec],
        )


class Disk(_Primitive):
    _op = _lib.OP_DISK

    def __init__(self, x0, r, rotate=[0, 0, 0], stretch=None, translate=None):
        self.xc, self.yc, self.r = x0[0], x0[1], r
        self._setup(2, (x0[0] - r, x0[0] + r, x0[1] - r, x0[1] + r), rotate, stretch, translate, None)
        self.corners = None

    def _params(self):
        return [self.xc, self.yc, self.r]

    def spec(self):
        return ("disk", dict(x0=[self.xc, self.yc], r=self.r, **self._tf_spec()))


class Ball(_Primitive):
    _op = _lib.OP_BALL

    def __init__(self, x0, r, rotate=[0, 0, 0], stretch=None, translate=None):
        if stretch is not None:
            assert len(stretch) == 3
        self.xc, self.yc, self.zc, self.r = x0[0], x0[1], x0[2], r
        box = (x0[0] - r, x0[0] + r, x0[1] - r, x0[1] + r, x0[2] - r, x0[2] + r)
        self._setup(3, box, rotate, stretch, translate, None)
        self.corners = None

    def _params(self):
        return [self.xc, self.yc, self.zc, self.r]

    def spec(self):
        return ("ball", dict(x0=[self.xc, self.yc, self.zc], r=self.r, **self._tf_spec()))


class Rectangle(_Primitive):
    _op = _lib.OP_RECT

    def __init__(self, bbox, rotate=[0.0, 0.0, 0.0], stretch=None, translate=None):
        self.bbox0 = bbox
        self._setup(2, bbox, rotate, stretch, translate, corners(bbox))

    def _params(self):
        return list(self.bbox0)

    def spec(self):
        return ("rectangle", dict(bbox=tuple(self.bbox0), **self._tf_spec()))


class Cube(_Primitive):
    _op = _lib.OP_CUBE

    def __init__(self, bbox, rotate=[0, 0, 0], stretch=None, translate=None):
        self.bbox0 = bbox
        self._setup(3, bbox, rotate, stretch, translate, corners(bbox))

    def _params(self):
        return list(self.bbox0)

    def spec(self):
        return ("cube", dict(bbox=tuple(self.bbox0), **self._tf_spec()))


class Torus(_Primitive):
    _op = _lib.OP_TORUS

    def __init__(self, r1, r2, rotate=[0, 0, 0], stretch=None, translate=None):
        """A torus with outer radius `r1` and inner radius `r2`."""
        assert r1 > 0.0 and r2 > 0.0
        z = 2 * max(r1, r2)
        self.t = (r1, r2)
        self._setup(3, (-2 * z, 2 * z, -2 * z, 2 * z, -2 * z, 2 * z), rotate, stretch, translate, None)
        self.corners = None

    def _params(self):
        return list(self.t)

    def spec(self):
        return ("torus", dict(r1=self.t[0], r2=self.t[1], **self._tf_spec()))


class Prism(_Primitive):
    _op = _lib.OP_PRISM

    def __init__(self, b, h, rotate=[0, 0, 0], stretch=None, translate=None):
        self.h = (b, h)
        self._setup(3, (-b, +b, -b, +b, -h, +h), rotate, stretch, translate, None)
        self.corners = None

    def _params(self):
        return list(self.h)

    def spec(self):
        return ("prism", dict(b=self.h[0], h=self.h[1], **self._tf_spec()))


class Cylinder(_Primitive):
    _op = _lib.OP_CYLINDER

    def __init__(self, h=1.0, r=0.5, rotate=[0, 0, 0], stretch=None, translate=None):
        assert h > 0.0 and r > 0.0
        self._h_full = h
        h = h / 2.0
        sz = max(h, r)
        self.h = (r, h)
        self._setup(3, (-2 * sz, 2 * sz, -2 * sz, 2 * sz, -2 * sz, 2 * sz), rotate, stretch, translate, None)
        self.corners = None

    def _params(self):
        return list(self.h)

    def spec(self):
        return ("cylinder", dict(h=self._h_full, r=self.h[0], **self._tf_spec()))


def _gather_corners(domains):
    cs = [d.corners for d in domains if d.corners is not None]
    return None if len(cs) == 0 else np.concatenate(cs)


def _lowerable(d):
    if not isinstance(d, _SDF):
        raise SDFProgramError(f"{type(d).__name__} is not a seismicmesh_b200 geometry object")
    return d


class _Combinator(_SDF):
    _hard = _smooth = 0
    _name = ""

    def __init__(self, domains, smoothness=0.0):
        dims = [d.dim for d in domains]
        assert all(x == dims[0] for x in dims), "all domains must have the same dimension"
        self.dim = dims[0]
        self.k = smoothness
        self.bbox = tuple(
            (min if k % 2 == 0 else max)(d.bbox[k] for d in domains) for k in range(2 * self.dim)
        )
        self.corners = _gather_corners(domains)
        self.domains = domains

    def _ordered(self):
        return list(self.domains)

    def _emit(self, out):
        kids = self._ordered()
        _lowerable(kids[0])._emit(out)
        for d in kids[1:]:
            _lowerable(d)._emit(out)
            rec = np.zeros(_lib.DM_SDF_WORDS)
            rec[0] = self._hard if self.k == 0.0 else self._smooth
            rec[2] = self.k
            out.append(rec)

    def spec(self):
        return (self._name, [d.spec() for d in self.domains], float(self.k))


class Union(_Combinator):
    _hard, _smooth, _name = _lib.OP_UNION, _lib.OP_SUNION, "union"


class Intersection(_Combinator):
    _hard, _smooth, _name = _lib.OP_INTER, _lib.OP_SINTER, "intersection"


class Difference(_Combinator):
    _hard, _smooth, _name = _lib.OP_DIFF, _lib.OP_SDIFF, "difference"

    def _ordered(self):
        # the smooth variant folds the reversed child list (reference :421-423)
        return list(self.domains) if self.k == 0.0 else list(self.domains)[::-1]


class Repeat(_SDF):
    def __init__(self, bbox, domain, period):
        self.bbox = bbox
        self.domain = domain
        self.corners = None
        self.period = np.array(period)
        self.parent = Cube(bbox)
        self.dim = 3

    def _emit(self, out):
        rec = np.zeros(_lib.DM_SDF_WORDS)
        rec[0] = _lib.OP_REPEAT_BEGIN
        rec[2:5] = self.period
        out.append(rec)
        _lowerable(self.domain)._emit(out)
        rec = np.zeros(_lib.DM_SDF_WORDS)
        rec[0] = _lib.OP_REPEAT_END
        rec[2:8] = self.bbox
        out.append(rec)

    def spec(self):
        return ("repeat", tuple(self.bbox), self.domain.spec(), [float(a) for a in self.period])


def lower(domain):
    """Device program of `domain`, or None if it is not a lowerable geometry object."""
    if isinstance(domain, _SDF):
        try:
            return domain.device_program()
        except SDFProgramError:
            return None
    return None


def from_spec(spec):
    """Build a geometry object from an oracle-style spec (inverse of ``.spec()``)."""
    kind = spec[0]
    if kind in ("union", "intersection", "difference"):
        cls = dict(union=Union, intersection=Intersection, difference=Difference)[kind]
        return cls([from_spec(c) for c in spec[1]], smoothness=spec[2])
    if kind == "repeat":
        return Repeat(tuple(spec[1]), from_spec(spec[2]), list(spec[3]))
    prm = dict(spec[1])
    kw = dict(rotate=list(prm.get("rotate") or [0.0, 0.0, 0.0]), stretch=prm.get("stretch"), translate=prm.get("translate"))
    if kind == "disk":
        return Disk(prm["x0"], prm["r"], **kw)
    if kind == "ball":
        return Ball(prm["x0"], prm["r"], **kw)
    if kind == "rectangle":
        return Rectangle(tuple(prm["bbox"]), **kw)
    if kind == "cube":
        return Cube(tuple(prm["bbox"]), **kw)
    if kind == "torus":
        return Torus(prm["r1"], prm["r2"], **kw)
    if kind == "prism":
        return Prism(prm["b"], prm["h"], **kw)
    if kind == "cylinder":
        return Cylinder(h=prm["h"], r=prm["r"], **kw)
    raise ValueError(kind)


# ---------------------------------------------------------------------------------------------
# The rest of the reference's `SeismicMesh.geometry` namespace that callers of the hot path use
# (geometry/__init__.py): the reference's own tests and benchmarks reach the mesh helpers through
# it (`SeismicMesh.geometry.simp_vol`, tests/test_2dmesher_SDF.py:29; `calc_dihedral_angles`,
# tests/test_3d_sliver.py:13).  Host helpers come from meshutil; the two natives of
# `_fast_geometry` that belong to the sliver loop run their CUDA kernels.
# ---------------------------------------------------------------------------------------------
from .meshutil import (  # noqa: E402,F401
    delete_boundary_entities,
    do_any_overlap,
    fix_mesh,
    get_boundary_edges,
    get_boundary_entities,
    get_boundary_facets,
    get_boundary_vertices,
    get_centroids,
    get_edges,
    get_facets,
    is_manifold,
    calc_re_ratios,
    get_winded_boundary_edges,
    laplacian2,
    laplacian2_fixed_point,
    linter,
    vertex_in_entity3,
    simp_qual,
    simp_vol,
    unique_rows,
    vertex_to_entities,
)


def calc_dihedral_angles(points, cells):
    """The six dihedral angles of every tetrahedron, (6T, 1) float64 cell-major, like
    `_fast_geometry.calc_dihedral_angles` (geometry/cpp/fast_geometry.cpp:351-452); device kernel."""
    p = D.points_dev(points, 3)
    t = D.to_dev(np.ascontiguousarray(cells), torch.int32)
    T = t.shape[0]
    out = torch.empty(6 * T, dtype=torch.float64, device=p.device)
    check(lib.dm_dihedral(D.ptr(p), D.ptr(t), T, 0.0, 4.0, D.ptr(out), None, D.stream_ptr()), "dm_dihedral")
    return out.cpu().numpy().reshape(-1, 1)


def calc_circumsphere_grad(p0, p1, p2, p3):
    """Gradient of the circumsphere radius with respect to p0, (S, 3), like
    `_fast_geometry.calc_circumsphere_grad` (fast_geometry.cpp:580-703); device kernel."""
    pts = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3) for a in (p0, p1, p2, p3)]
    S = len(pts[0])
    p = D.points_dev(np.concatenate(pts, axis=0), 3)
    t = D.to_dev(np.arange(4 * S, dtype=np.int32).reshape(4, S).T.copy(), torch.int32)
    out = torch.empty((S, 3), dtype=torch.float64, device=p.device)
    check(lib.dm_circumsphere_grad(D.ptr(p), D.ptr(t), None, S, D.ptr(out), D.stream_ptr()), "dm_circumsphere_grad")
    return out.cpu().numpy()


def unique_edges(edges):
    """Sorted unique (min, max) vertex pairs of an (K, 2) integer array, like
    `_fast_geometry.unique_edges` (fast_geometry.cpp:49-77).  The device pipeline builds bars from
    CELLS (engine.unique_bars, stage B); a bare edge list is handled as degenerate triangles (v, w, w)."""
    from .engine import unique_bars

    e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 2)
    if len(e) == 0:
        return np.zeros((0, 2), dtype=np.int32)
    tri = np.column_stack([e[:, 0], e[:, 1], e[:, 1]]).astype(np.int32)
    out = unique_bars(tri)
    # the degenerate triangle (v, w, w) also carries the pair (w, w): keep it only where the input has it
    loops = np.unique(e[e[:, 0] == e[:, 1], 0])
    keep = (out[:, 0] != out[:, 1]) | np.isin(out[:, 0], loops)
    return np.ascontiguousarray(out[keep])


__all__ += [
    "simp_vol", "simp_qual", "fix_mesh", "get_edges", "get_facets", "get_centroids", "get_boundary_edges",
    "get_boundary_facets", "get_boundary_vertices", "get_boundary_entities", "delete_boundary_entities",
    "laplacian2_fixed_point", "linter", "do_any_overlap", "is_manifold", "vertex_to_entities", "unique_rows",
    "calc_dihedral_angles", "calc_circumsphere_grad", "unique_edges", "calc_re_ratios", "get_winded_boundary_edges",
    "laplacian2", "vertex_in_entity3",
]
