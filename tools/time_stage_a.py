#!/usr/bin/env python
"""Developer utility: time stage A (prep + cull_scatter) alone on the bench workload, e.g. for an
experimental build selected with DM_LIB_PATH.  Prints mean ms over --reps launches (L2 flushed)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seismicmesh_b200 as sm  # noqa: E402
from bench import make_points, triangulate  # noqa: E402
from seismicmesh_b200 import device as D  # noqa: E402
from seismicmesh_b200._lib import check, lib  # noqa: E402
from seismicmesh_b200.engine import ForceLoop, Level, SizeSpec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="ball")
ap.add_argument("--h0", type=float, default=0.02)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--triangulator", default=None, choices=[None, "native", "qhull"],
                help="whose cells (and cell ORDER: stage A is sensitive to it) to feed; default: the package's")
ap.add_argument("--shuffle", action="store_true", help="feed the cells in random order (worst case for stage A)")
a = ap.parse_args()
p, dim = make_points(a.workload, a.h0)
if a.triangulator is None:
    t, _ = triangulate(p)
else:
    from seismicmesh_b200.triangulator import get_triangulator  # noqa: E402

    t = get_triangulator(a.triangulator, dim).triangulate(p)
if a.shuffle:
    t = np.ascontiguousarray(t[np.random.default_rng(0).permutation(len(t))])
dom = sm.Ball([0.0, 0.0, 0.0], 1.0) if dim == 3 else sm.Disk([0.0, 0.0], 1.0)
loop = ForceLoop(dim, [Level(dom, dim)], SizeSpec(dim, const=a.h0), a.h0, 0.1 * a.h0, 1e-8 * a.h0)
pd, td = D.to_dev(p, torch.float64), D.to_dev(t, torch.int32)
plan = loop.ensure_plan(len(p), len(t))
import ctypes as C  # noqa: E402

flush = torch.empty(512 << 20, dtype=torch.uint8, device=pd.device)
ms = []
for r in range(a.reps + 3):
    flush.fill_(r & 0xFF)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib.dm_stage_cull_count(C.byref(plan.c), D.ptr(loop._progs[0]), D.ptr(pd), D.ptr(td), 0.1 * a.h0, 1, D.stream_ptr()), "stage A")
    e1.record()
    torch.cuda.synchronize()
    if r >= 3:
        ms.append(e0.elapsed_time(e1))
print(f"stage A ({os.environ.get('DM_LIB_PATH', 'default')}, cells: {a.triangulator or 'package default'}{', shuffled' if a.shuffle else ''}): mean {np.mean(ms):.4f} ms  min {np.min(ms):.4f} ms")
