#!/bin/bash
# h2d peak + GPU suite + default bench line (threaded host Delaunay in the time-to-mesh block)
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2ce}
python tools/gpu/h2d_peak.py gpurun_out/${TAG}_h2d_peak.json
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_b.err; echo "bench rc=$?"
TAG=$TAG python - <<'PY'
import json,os
d=json.loads(open("gpurun_out/%s_b.json"%os.environ["TAG"]).read().strip().splitlines()[-1])
print("ms/step",d["ms_per_step"],"e2e",d["e2e"]["ms_per_step"],"delaunay_s",d["delaunay_s"],d["delaunay_backend"])
for w,v in d['workloads'].items(): print(w, "delaunay_s", v['delaunay_s'], "ms", v['ms_per_step'])
for k,v in (d.get("time_to_mesh") or {}).items():
    for kk,vv in v.items():
        if isinstance(vv,dict): print(k,kk,"wall",round(vv["wall_s"],3),"delaunay",round(vv["delaunay_s"],3),"tri",vv["triangulations"],"q",round(vv["mean_quality"],4),round(vv["min_quality"],4),vv["vertices"],vv["cells"])
PY
