"""Thread scaling of the host 3-D triangulator (dmh_delaunay3d_mt) on DistMesh-shaped input: the
reference's staggered lattice clipped to the unit ball, jittered by 0.1 h0 (general position).
Usage: python tools/host/time_delaunay3d_threads.py [out.json]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from seismicmesh_b200.generation import _staggered_grid  # noqa: E402
from seismicmesh_b200.triangulator import BowyerWatsonTriangulator  # noqa: E402

out = {"cores": len(os.sched_getaffinity(0)), "runs": []}
for h0 in (0.05, 0.02):
    s = np.ascontiguousarray(_staggered_grid(h0, 3, np.array([[-1.0, 1.0]] * 3)))
    p = s[(s ** 2).sum(1) < 1.0]
    p = np.ascontiguousarray(p + np.random.default_rng(0).uniform(-0.1 * h0, 0.1 * h0, p.shape))
    base = None
    for th in (1, 2, 4, 8, 12, 16, 24, 31):
        if th > out["cores"]:
            break
        tri = BowyerWatsonTriangulator(3, threads=th)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            t = tri.triangulate(p)
            best = min(best, time.perf_counter() - t0)
        base = best if base is None else base
        rec = {"h0": h0, "N": len(p), "T": len(t), "threads": th, "s": round(best, 4), "speedup": round(base / best, 2)}
        out["runs"].append(rec)
        print(rec, flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
