#!/bin/bash
# 2-GPU slab line with the threaded host Delaunay (threads = cores / ranks) + the 2-rank GPU tests
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2cj}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${TAG}_b2.json 2> gpurun_out/${TAG}_b2.err; echo "bench2 rc=$?"
tail -c 600 gpurun_out/${TAG}_b2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_b2.json") if l.startswith("{")][-1])
print("2 GPUs: ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d.get("delaunay_backend"), "delaunay_s", d.get("delaunay_s"))
for k,v in (d.get("workloads") or {}).items(): print(k, v.get("ms_per_step"), v.get("value"))
PY
