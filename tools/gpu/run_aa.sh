#!/bin/bash
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2ah}
timeout 900 python -m pytest tests -m gpu -x -q -k "${K:-pad or sizing or segy or gridded or parallel}" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
python - <<'PY'
import time, sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
import seismicmesh_b200 as sm
for w in ("bp2004", "eage"):
    vp, bbox = bench.synth_vp(w)
    hmin, fr, dim, kw = bench.sizing_kwargs(w, vp)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ef = sm.get_sizing_function_from_segy(None, bbox, velocity_data=vp.copy(), **kw)
        f = ef.interpolant().struct(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(w, "sizing + interpolant on device", round(t1 - t0, 3), "s  grid", ef.interpolant().shape)
PY
