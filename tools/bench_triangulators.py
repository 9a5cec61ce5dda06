#!/usr/bin/env python
"""Host-only: time the native triangulators against Qhull on the bench's own point sets and check the
cells are the same set (points in general position).  Prints one JSON object; no GPU needed beyond
importing the package.  `python tools/bench_triangulators.py > profiles/<round>_host_triangulators.json`"""
import json
import os
import platform
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_points  # noqa: E402
from seismicmesh_b200.triangulator import get_triangulator  # noqa: E402


def canon(t):
    t = np.sort(np.asarray(t, dtype=np.int64), axis=1)
    return t[np.lexsort(t.T[::-1])]


def best(f, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = f()
        ts.append(time.perf_counter() - t0)
    return out, min(ts)


rows = []
for workload, h0 in (("disk", 0.01), ("disk", 0.002), ("ball", 0.05), ("ball", 0.03), ("ball", 0.02)):
    p, dim = make_points(workload, h0)
    native, qhull = get_triangulator("native", dim), get_triangulator("qhull", dim)
    reps = 3 if len(p) < 200000 else 1
    tn, sn = best(lambda: native.triangulate(p), reps)
    tq, sq = best(lambda: qhull.triangulate(p), reps)
    rows.append({"points": f"{workload} h0={h0}", "N": int(len(p)), "cells": int(len(tn)), "native_s": round(sn, 4),
                 "qhull_s": round(sq, 4), "speedup": round(sq / sn, 2), "same_cell_set": bool(np.array_equal(canon(tn), canon(tq))),
                 "native": native.name, "qhull_takeovers": int(native.qhull_retries)})
    print(rows[-1], file=sys.stderr, flush=True)
print(json.dumps({"what": "host Delaunay of the bench point sets, one core, best of 3 (1 above 200 k points)",
                  "host": platform.processor() or platform.machine(), "cores_visible": os.cpu_count(), "rows": rows}, indent=1))
