#!/usr/bin/env python
"""Developer diagnostic for the reference's test_pfix scenario: which fixed point ends up without a
cell at termination, and why (cells around it in the last triangulation, fd at their centroids)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seismicmesh_b200 as sm  # noqa: E402
from seismicmesh_b200 import generation  # noqa: E402

hmin = 0.05
bbox = (0.0, 1.0, 0.0, 1.0, 0.0, 1.0)
pfix = np.linspace((0.0, 0.0, 0.0), (1.0, 0.0, 1.0), int(np.sqrt(2) / hmin))
pfix = np.vstack((pfix, sm.geometry.corners(bbox)))
cube = sm.Cube(bbox)
cap = {}
orig = generation._termination


def spy(p, t, opts, dim, **kw):
    cap["p"], cap["t"] = p.copy(), t.copy()
    return orig(p, t, opts, dim, **kw)


generation._termination = spy
points, cells = sm.generate_mesh(domain=cube, edge_length=hmin, pfix=pfix, verbose=0)
p, t = cap["p"], cap["t"]
print("N at termination", len(p), "kept cells", len(t), "final", len(points), len(cells), sm.last_run_stats)
allpf = np.vstack((pfix, sm.geometry.corners(bbox)))  # rows 0..nfix-1 of p (domain corners appended by generate_mesh)
print("rows 0..nfix-1 unchanged:", np.abs(p[: len(allpf)] - allpf).max())
from scipy.spatial import Delaunay  # noqa: E402

full = Delaunay(p).simplices
geps = 0.1 * hmin
for q in pfix:
    d2 = ((points - q) ** 2).sum(1)
    if d2.min() > 1e-12:
        rows = np.flatnonzero(((p - q) ** 2).sum(1) < 1e-20)
        print("MISSING fixed point", q, "rows in p:", rows, "nearest final vertex at", np.sqrt(d2.min()))
        for r in rows:
            inc = full[(full == r).any(1)]
            kept = t[(t == r).any(1)]
            print("  row", r, "cells in full Delaunay:", len(inc), "kept:", len(kept))
            for c in inc:
                cen = p[c].sum(0) / 4
                print("    cell", c.tolist(), "fd(centroid)", float(cube.eval(cen[None])[0]), "threshold", -geps,
                      "nbr dists", np.round(np.sqrt(((p[c] - q) ** 2).sum(1)), 4).tolist())
