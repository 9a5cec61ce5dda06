#!/bin/bash
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2ab}
timeout 600 python -m pytest tests -m gpu -x -q -k "${K:-laplacian}" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${TAG}_pytest.log
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, ".")
import bench, torch
import seismicmesh_b200 as sm
from seismicmesh_b200 import meshutil as mu
from seismicmesh_b200.engine import laplacian_smooth
from scipy.spatial import Delaunay
for h0 in (0.01, 0.004):
    p, dim = bench.make_points("disk", h0)
    t = Delaunay(p).simplices.astype(np.int32)
    c = p[t].sum(1) / 3
    t = t[np.hypot(c[:, 0], c[:, 1]) - 1 < -0.001]
    p, t, _ = mu.fix_mesh(p, t, dim=2, delete_unused=True)
    laplacian_smooth(p.copy(), t.copy())
    t0 = time.perf_counter(); got, _ = laplacian_smooth(p.copy(), t.copy()); torch.cuda.synchronize(); t1 = time.perf_counter()
    ref, _ = mu.laplacian2_fixed_point(p.copy(), t.copy()); t2 = time.perf_counter()
    print(f"disk h0={h0}: N={len(p)} device {t1-t0:.4f}s iters={laplacian_smooth.last} host LU {t2-t1:.4f}s maxdiff {np.abs(got-ref).max():.2e}")
PY
