"""ttol sweep (time-to-mesh): Delaunay calls / wall time / quality against the reference semantics."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import seismicmesh_b200 as sm
from seismicmesh_b200 import meshutil

sys.path.insert(0, ".")
import bench

out = {}
cases = [("ball_h0=0.05", sm.Ball([0.0, 0.0, 0.0], 1.0), 0.05, 3)]
vp, bbox = bench.synth_vp("eage")
hmin, fr, dim, kw = bench.sizing_kwargs("eage", vp, 150.0, 2.0)
ef = sm.get_sizing_function_from_segy(None, bbox, velocity_data=vp, **kw)
cases.append(("eage_hmin150", sm.Cube(ef.bbox), ef, 3))
vp, bbox = bench.synth_vp("bp2004")
hmin, fr, dim, kw = bench.sizing_kwargs("bp2004", vp, 75.0, 2.0)
ef2 = sm.get_sizing_function_from_segy(None, bbox, velocity_data=vp, **kw)
cases.append(("bp2004_hmin75", sm.Rectangle(ef2.bbox), ef2, 2))
for name, dom, edge, dim in cases:
    for iters in (25, 50):
        for ttol in (None, 0.1, 0.2, 0.3, 0.5):
            kw = {} if ttol is None else {"ttol": ttol}
            c0 = time.perf_counter()
            p, t = sm.generate_mesh(dom, edge, max_iter=iters, verbose=0, **kw)
            w = time.perf_counter() - c0
            st = dict(sm.last_run_stats)
            q = meshutil.simp_qual(p, t)
            rec = dict(wall_s=round(w, 3), tri=st["triangulations"], delaunay_s=round(st["delaunay"], 3), nv=len(p), nc=len(t),
                       mean_q=round(float(q.mean()), 5), min_q=round(float(q.min()), 4))
            if dim == 3:
                dh = sm.geometry.calc_dihedral_angles(p, t)
                rec["slivers"] = int(((dh.reshape(-1, 6) < 10 * np.pi / 180).any(axis=1)).sum())
            out[f"{name}|{iters}|{ttol}"] = rec
            print(name, iters, ttol, rec, flush=True)
json.dump(out, open("gpurun_out/r2g_ttol_sweep.json", "w"), indent=1)
