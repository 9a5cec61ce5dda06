#!/bin/bash
# 2 GPUs: slab bench through the package's parallel path (ball default, then eage) + the parallel GPU tests
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2ac}
NG=${NG:-2}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/${TAG}_bench${NG}.json 2> gpurun_out/${TAG}_bench${NG}.err ) 2> gpurun_out/${TAG}_time.txt
tail -12 gpurun_out/${TAG}_bench${NG}.err; cat gpurun_out/${TAG}_time.txt
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench${NG}.json") if l.startswith("{")][-1])
    print("${NG}gpu", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print(d["config"]["parallelism"]); print(d["config"].get("stage_ab_layout"))
    w = d.get("workloads") or {}
    for k, v in w.items(): print(k, {kk: v.get(kk) for kk in ("ms_per_step", "value")} if isinstance(v, dict) else v)
except Exception as e:
    print("ERR", e)
PY
