"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (krober10nd/SeismicMesh,
mounted read-only at /root/reference) so that golden vectors can be generated from the
reference's own functions.  Never imported by the product package.

The reference needs CGAL, mpi4py, matplotlib, pyamg, segyio and h5py; none of them exist in
this image.  Everything on the DistMesh hot path is CGAL-free, so we provide ``sys.modules``
stand-ins for what is missing:

* ``mpi4py.MPI``            -> a rank-0-of-1 ``COMM_WORLD``
* ``matplotlib``/``pyamg``/``_delaunay``/``_cpputils`` -> empty modules (+ a sparse-LU
  ``ruge_stuben_solver`` so ``laplacian2_fixed_point`` still solves its linear system)
* ``_delaunay_class(3)``    -> ``scipy.spatial.Delaunay`` (Qhull) behind the reference's
  ``insert/move/get_finite_vertices/get_finite_cells`` interface.  Vertex order is kept.
* ``segyio``                -> a plain struct-based SEG-Y trace decoder (``_SegyFile``), so the reference's
  own tests on ``tests/testing.segy`` can be replayed
* ``_fast_geometry``/``_FastHJ`` -> the reference's OWN C++ sources compiled by
  ``oracle/Makefile`` into ``oracle/_ref`` (not copied).

and patch ``geometry.utils.unique_rows`` which is not NumPy-2 clean
(``geometry/utils.py:157-158`` uses ``np.character`` and a 2-D ``return_inverse``).

This harness only works where /root/reference exists (the build container).  On the GPU box
only the committed fixtures under tests/golden/ and the compiled ``oracle/_ref/*.so`` exist.
"""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SEISMICMESH_REFERENCE", "/root/reference")
REF_NATIVE = os.path.join(HERE, "_ref")

_loaded = None


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "SeismicMesh"))


def native_available():
    if not os.path.isdir(REF_NATIVE):
        return False
    return any(f.startswith("_fast_geometry") for f in os.listdir(REF_NATIVE))


def load_native():
    """Import the reference's own compiled `_fast_geometry` (oracle/_ref). Travels to GPU box."""
    if not native_available():
        raise ImportError("oracle/_ref/_fast_geometry*.so missing: run `make -C oracle`")
    if REF_NATIVE not in sys.path:
        sys.path.insert(0, REF_NATIVE)
    return importlib.import_module("_fast_geometry")


def load_native_fasthj():
    """Import the reference's own compiled `_FastHJ` (oracle/_ref): the gradient limiter."""
    if not native_available() or not any(f.startswith("_FastHJ") for f in os.listdir(REF_NATIVE)):
        raise ImportError("oracle/_ref/_FastHJ*.so missing: run `make -C oracle`")
    if REF_NATIVE not in sys.path:
        sys.path.insert(0, REF_NATIVE)
    return importlib.import_module("_FastHJ")


class _FakeComm:
    rank = 0
    size = 1

    def bcast(self, x, root=0):
        return x

    def allreduce(self, x, op=None):
        return x

    def barrier(self):
        pass


def _qhull(points, dim):
    from scipy.spatial import Delaunay

    return Delaunay(points).simplices.astype(np.int32)


class _DT:
    dim = 2

    def __init__(self):
        self.p = np.zeros((0, self.dim))

    def insert(self, flat):
        q = np.asarray(flat, dtype=np.float64).reshape(-1, self.dim)
        self.p = np.vstack((self.p, q))

    def move(self, idx, flat):
        q = np.asarray(flat, dtype=np.float64).reshape(-1, self.dim)
        self.p[np.asarray(idx, dtype=np.int64)] = q

    def get_finite_vertices(self):
        return self.p.copy()

    def get_finite_cells(self):
        return _qhull(self.p, self.dim)


class _DT2(_DT):
    dim = 2


class _DT3(_DT):
    dim = 3


class _SegyFile:
    """Stand-in for ``segyio.open(filename, ignore_geometry=True)`` as the reference uses it
    (sizing/mesh_size_function.py:633-646: ``len(f.samples)``, ``len(f.trace)``, iteration over
    ``f.trace``).  A deliberately plain, value-by-value decoder (struct), independent of the product's
    vectorised reader, for IBM-float (format 1) and IEEE-float (format 5) files."""

    def __init__(self, filename):
        import struct

        with open(filename, "rb") as f:
            raw = f.read()
        ns = struct.unpack(">H", raw[3220:3222])[0]
        fmt = struct.unpack(">H", raw[3224:3226])[0]
        if fmt not in (1, 5):
            raise NotImplementedError(f"SEG-Y format code {fmt}")
        stride = 240 + 4 * ns
        ntr = (len(raw) - 3600) // stride

        def ibm(w):
            sign = -1.0 if w >> 31 else 1.0
            return sign * ((w & 0xFFFFFF) / float(1 << 24)) * 16.0 ** (((w >> 24) & 0x7F) - 64)

        self.samples = list(range(ns))
        self.trace = []
        for k in range(ntr):
            off = 3600 + k * stride + 240
            if fmt == 1:
                words = struct.unpack(f">{ns}I", raw[off:off + 4 * ns])
                self.trace.append(np.array([ibm(w) for w in words], dtype=np.float32))
            else:
                self.trace.append(np.array(struct.unpack(f">{ns}f", raw[off:off + 4 * ns]), dtype=np.float32))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _unique_rows_np2(A, return_index=False, return_inverse=False):
    A = np.require(A, requirements="C")
    assert A.ndim == 2
    view = A.view(np.dtype((np.void, A.dtype.itemsize * A.shape[1])))
    B, I, J = np.unique(view, return_index=True, return_inverse=True)
    B = B.view(A.dtype).reshape((-1, A.shape[1]), order="C")
    J = J.ravel()
    if return_index:
        return (B, I, J) if return_inverse else (B, I)
    return (B, J) if return_inverse else B


def load_reference():
    """Return the reference `SeismicMesh` package, imported with the shims above."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise ImportError(f"reference not found under {REF_ROOT}")
    load_native()

    mpi4py = types.ModuleType("mpi4py")
    MPI = types.ModuleType("mpi4py.MPI")
    MPI.COMM_WORLD = _FakeComm()
    MPI.Intracomm = _FakeComm
    MPI.MIN = "min"
    MPI.SUM = "sum"
    mpi4py.MPI = MPI
    sys.modules.setdefault("mpi4py", mpi4py)
    sys.modules.setdefault("mpi4py.MPI", MPI)

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)

    pyamg = types.ModuleType("pyamg")

    class _LU:
        def __init__(self, A):
            from scipy.sparse.linalg import splu

            self.lu = splu(A.tocsc().astype(np.float64))

        def solve(self, b):
            return self.lu.solve(np.asarray(b, dtype=np.float64))

    pyamg.ruge_stuben_solver = lambda A: _LU(A)
    sys.modules.setdefault("pyamg", pyamg)
    for name in ("_cpputils", "h5py"):
        sys.modules.setdefault(name, types.ModuleType(name))
    segyio = types.ModuleType("segyio")
    segyio.open = lambda filename, ignore_geometry=True: _SegyFile(filename)
    sys.modules.setdefault("segyio", segyio)
    dl = types.ModuleType("_delaunay")

    def _not_available(*a, **k):
        raise NotImplementedError("CGAL-only helper, not on the hot path")

    dl._circumballs2 = dl._circumballs3 = _not_available
    dl._delaunay2 = lambda x, y: _qhull(np.column_stack((x, y)), 2)
    dl._delaunay3 = lambda x, y, z: _qhull(np.column_stack((x, y, z)), 3)
    sys.modules.setdefault("_delaunay", dl)
    d2 = types.ModuleType("_delaunay_class")
    d2.DelaunayTriangulation = _DT2
    d3 = types.ModuleType("_delaunay_class3")
    d3.DelaunayTriangulation3 = _DT3
    sys.modules.setdefault("_delaunay_class", d2)
    sys.modules.setdefault("_delaunay_class3", d3)

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    sm = importlib.import_module("SeismicMesh")
    sm.geometry.utils.unique_rows = _unique_rows_np2
    sm.geometry.unique_rows = _unique_rows_np2
    _loaded = sm
    return sm
