#!/bin/bash
# round 2, call F: the default bench invocation (headline + sub-records + time-to-mesh)
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
( time timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err ) 2> gpurun_out/r2f_time.txt
tail -30 gpurun_out/r2f_bench.err; cat gpurun_out/r2f_time.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2f_bench.json"))
print("headline", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("cpu", d["cpu_baseline"])
for k, v in (d.get("workloads") or {}).items():
    print(k, v["ms_per_step"], v["value"], "e2e", v["e2e"]["value"], "whole", v["roofline"]["whole_step"]["frac"], "dp", v.get("max_abs_dp_vs_oracle"))
print(json.dumps(d.get("time_to_mesh"), indent=1))
PY
