"""Where the sizing preprocessing spends its time on the EAGE-shaped model (run under gpurun)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
import seismicmesh_b200 as sm
from seismicmesh_b200 import sizing as S

for w in ("eage", "bp2004"):
    vp, bbox = bench.synth_vp(w)
    hmin, fr, dim, kw = bench.sizing_kwargs(w, vp)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ef = sm.get_sizing_function_from_segy(None, bbox, velocity_data=vp.copy(), **kw)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ef.interpolant().struct()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    print(w, "grid", vp.shape, "sizing %.3f s (limiter sweeps %d)" % (t1 - t0, S.limgrad.last_sweeps), "interpolant/corner records %.3f s" % (t2 - t1))
    a = torch.from_numpy(vp).cuda()
    torch.cuda.synchronize(); t0 = time.perf_counter(); b = torch.from_numpy(vp.copy()).cuda(); torch.cuda.synchronize()
    print("   pageable H2D of vp %.3f s; np.where scan + copy on the host:" % (time.perf_counter() - t0), end=" ")
    t0 = time.perf_counter(); v2 = vp.copy(); pos = np.where(v2 < 1e-3); print("%.3f s" % (time.perf_counter() - t0))
