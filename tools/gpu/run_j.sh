#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q -k "sizing or limgrad or segy or gridded" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -12 gpurun_out/r2j_pytest.log
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import bench, seismicmesh_b200 as sm
for wk, h, f in (("bp2004", 75.0, 2.0), ("eage", 150.0, 2.0)):
    vp, bbox = bench.synth_vp(wk)
    hmin, fr, dim, kw = bench.sizing_kwargs(wk, vp, h, f)
    for rep in range(2):
        t0 = time.perf_counter()
        ef = sm.get_sizing_function_from_segy(None, bbox, velocity_data=vp.copy(), **kw)
        print(wk, "sizing_s", round(time.perf_counter() - t0, 3), flush=True)
PY
