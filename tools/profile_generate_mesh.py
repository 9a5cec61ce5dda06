#!/usr/bin/env python
"""Developer utility: where does the wall time of one generate_mesh call go?  Prints
`last_run_stats` (host Delaunay / H2D / device / D2H / termination) and the cProfile top entries."""
import argparse
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seismicmesh_b200 as sm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--h0", type=float, default=0.01)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--dim", type=int, default=2)
ap.add_argument("--ttol", type=float, default=None)
a = ap.parse_args()
dom = sm.Disk([0.0, 0.0], 1.0) if a.dim == 2 else sm.Ball([0.0, 0.0, 0.0], 1.0)
kw = {} if a.ttol is None else {"ttol": a.ttol}
sm.generate_mesh(dom, a.h0 * 4, max_iter=3, verbose=0)  # warm-up (CUDA context, library load)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
p, t = sm.generate_mesh(dom, a.h0, max_iter=a.iters, verbose=0, **kw)
pr.disable()
print(f"wall {time.perf_counter() - t0:.3f} s  N={len(p)} T={len(t)}")
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in sm.last_run_stats.items()})
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
