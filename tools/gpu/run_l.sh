#!/bin/bash
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -5 gpurun_out/r2l_pytest.log
for w in "eage 75 4" "bp2004 25 6" "eage 150 2" "bp2004 75 2"; do
  set -- $w
  timeout 300 python bench.py --workload $1 --h0 $2 --freq $3 --steps 10 --warmup 3 --no-extras --kernel-table gpurun_out/r2l_kernels_$1_$2.json > gpurun_out/r2l_bench_$1_$2.json 2> gpurun_out/r2l_bench_$1_$2.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2l_kernels_$1_$2.json"))
b = json.loads([l for l in open("gpurun_out/r2l_bench_$1_$2.json") if l.startswith("{")][-1])
print("$1 $2", [(k["kernel"], round(k["ms"], 4)) for k in d["kernels"]], "ms/step", round(b["ms_per_step"], 4), "dp", b["cpu_baseline"]["max_abs_dp_vs_oracle"])
PY
done
