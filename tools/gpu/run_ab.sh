#!/bin/bash
# threaded host Delaunay on the GPU box's cores + GPU suite + default bench line
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2ca}
(nproc; lscpu | grep -i "model name\|^CPU(s)\|thread(s) per\|socket\|numa node") > gpurun_out/${TAG}_host.txt 2>&1
timeout 600 python tools/host/time_delaunay3d_threads.py gpurun_out/${TAG}_delaunay3d_threads.json > gpurun_out/${TAG}_delaunay3d_threads.log 2>&1
tail -20 gpurun_out/${TAG}_delaunay3d_threads.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_b.err; echo "bench rc=$?"
python - <<'PY'
import json,os
d=json.loads(open(os.environ.get("TAGF","gpurun_out/%s_b.json"%os.environ.get("TAG","r2ca"))).read().strip().splitlines()[-1])
print("ms/step",d["ms_per_step"],"e2e",d["e2e"]["ms_per_step"],"delaunay_s",d["delaunay_s"],d["delaunay_backend"])
for k,v in (d.get("time_to_mesh") or {}).items():
    for kk,vv in v.items():
        if isinstance(vv,dict): print(k,kk,"wall",round(vv["wall_s"],3),"delaunay",round(vv["delaunay_s"],3),"tri",vv["triangulations"])
PY
