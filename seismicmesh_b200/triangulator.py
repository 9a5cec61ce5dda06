"""Host Delaunay retriangulation -- deliberately NOT a GPU kernel (BASELINE.json north_star:
"Delaunay retriangulation stays on CGAL on the host ... its time is reported separately").

The reference wraps CGAL's ``Delaunay_triangulation_2/3`` (generation/cpp/delaunay_class*.cpp) and
renumbers the vertices on every call; here the triangulator sits behind one small interface that
keeps the input vertex order, so device-resident coordinates never have to be permuted or
re-uploaded.  CGAL is not installed in this image, so the backend is Qhull through
``scipy.spatial.Delaunay`` (same bar set for points in general position; tie-breaking on the
co-circular initial lattice differs).
"""
import numpy as np


class QhullTriangulator:
    name = "qhull (scipy.spatial.Delaunay)"

    def __init__(self, dim):
        self.dim = dim

    def triangulate(self, points):
        """points (N,dim) float64 host array -> cells (T,dim+1) int32, vertex ids = input rows."""
        from scipy.spatial import Delaunay

        return np.ascontiguousarray(Delaunay(points).simplices, dtype=np.int32)


def get_triangulator(spec, dim):
    if spec is None or spec == "qhull":
        return QhullTriangulator(dim)
    if hasattr(spec, "triangulate"):
        return spec
    raise ValueError(f"unknown triangulator {spec!r}")
