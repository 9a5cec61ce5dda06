#!/bin/bash
# round 2, call H (2 GPUs): slab bench through the package's parallel path + gpu parallel tests
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2h_bench2.json 2> gpurun_out/r2h_bench2.err ) 2> gpurun_out/r2h_time.txt
tail -25 gpurun_out/r2h_bench2.err; cat gpurun_out/r2h_time.txt
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2h_bench2.json"))
    print("2gpu", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print(d["config"]["parallelism"])
    print(json.dumps(d.get("workloads"), indent=1)[:3000])
except Exception as e:
    print("ERR", e)
PY
