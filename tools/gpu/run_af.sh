#!/bin/bash
# N-GPU slab line + the N > 1 time-to-mesh block (NG ranks)
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2cm}; NG=${NG:-4}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/${TAG}_b${NG}.json 2> gpurun_out/${TAG}_b${NG}.err ) 2> gpurun_out/${TAG}_time${NG}.txt; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_time${NG}.txt
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_b${NG}.json") if l.startswith("{")][-1])
print("${NG} GPUs: ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d.get("delaunay_backend"))
for k,v in (d.get("workloads") or {}).items(): print(k, v.get("ms_per_step"), v.get("value"), v.get("error"))
print(json.dumps(d["time_to_mesh"],indent=1))
PY
