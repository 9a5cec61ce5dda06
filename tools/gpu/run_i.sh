#!/bin/bash
# suite + default bench at HEAD
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -12 gpurun_out/r2i_pytest.log
( time timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err ) 2> gpurun_out/r2i_time.txt
grep -v "^  " gpurun_out/r2i_bench.err | tail; grep -A6 "eage\|bp2004" gpurun_out/r2i_bench.err | head -30; cat gpurun_out/r2i_time.txt
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2i_bench.json") if l.startswith("{")][-1])
print("headline", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
for k, v in (d.get("workloads") or {}).items():
    print(k, v["ms_per_step"], v["value"], "e2e", v["e2e"]["value"], "whole", v["roofline"]["whole_step"]["frac"], "dp", v.get("max_abs_dp_vs_oracle"))
for k, v in (d.get("time_to_mesh") or {}).items():
    print(k, {m: (r["wall_s"], r["triangulations"], round(r["mean_quality"], 4)) for m, r in v.items() if isinstance(r, dict)})
PY
