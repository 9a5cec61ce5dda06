#!/bin/bash
# GPU suite + default bench line + kernel tables of the six workloads (no ncu, no sanitizers)
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2cr}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
( time timeout 1500 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err ) 2> gpurun_out/${TAG}_bench_default_time.txt; tail -3 gpurun_out/${TAG}_bench_default_time.txt
for w in "ball 0.02 0" "disk 0.01 0" "eage 150 2" "bp2004 75 2" "eage 75 4" "bp2004 25 6"; do
  set -- $w
  F=""; if [ "$3" != "0" ]; then F="--freq $3"; fi
  timeout 400 python bench.py --workload $1 --h0 $2 $F --steps 20 --warmup 3 --no-extras --kernel-table gpurun_out/${TAG}_kernels_$1_$2.json > gpurun_out/${TAG}_bench_$1_$2.json 2> gpurun_out/${TAG}_bench_$1_$2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_kernels_$1_$2.json"))
    b = json.loads([l for l in open("gpurun_out/${TAG}_bench_$1_$2.json") if l.startswith("{")][-1])
    print("$1 $2", [(k["kernel"][:10], round(k["ms"], 4)) for k in d["kernels"]], "ms/step", round(b["ms_per_step"], 4), "dp", b["cpu_baseline"]["max_abs_dp_vs_oracle"], "launches", b["gpu_launches"])
except Exception as e:
    print("$1 $2 ERR", e)
PY
done
