#!/bin/bash
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -5 gpurun_out/r2o_pytest.log
for w in "ball 0.02 0" "disk 0.01 0" "eage 150 2" "bp2004 75 2"; do
  set -- $w
  F=""; if [ "$3" != "0" ]; then F="--freq $3"; fi
  timeout 300 python bench.py --workload $1 --h0 $2 $F --steps 20 --warmup 3 --no-extras --kernel-table gpurun_out/r2o_kernels_$1.json > gpurun_out/r2o_bench_$1.json 2> gpurun_out/r2o_bench_$1.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2o_kernels_$1.json"))
b = json.loads([l for l in open("gpurun_out/r2o_bench_$1.json") if l.startswith("{")][-1])
print("$1 $2", [(k["kernel"][:10], round(k["ms"], 4)) for k in d["kernels"]], "ms/step", round(b["ms_per_step"], 4), "dp", b["cpu_baseline"]["max_abs_dp_vs_oracle"], "launches", b["gpu_launches"])
PY
done
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_memcheck.log 2>&1; tail -3 gpurun_out/r2o_memcheck.log
