#!/bin/bash
# round 2, call R: first run of the tile layout (dm_tiles.cuh): smoke, GPU suite, kernel tables tiles vs buckets
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2r_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -15 gpurun_out/r2r_pytest.log
for w in "ball 0.02 0" "eage 150 2" "disk 0.01 0" "bp2004 25 6"; do
  set -- $w
  F=""; if [ "$3" != "0" ]; then F="--freq $3"; fi
  for v in 1 0; do
    export DM_TILES=$v
    timeout 300 python bench.py --workload $1 --h0 $2 $F --steps 20 --warmup 3 --no-extras --kernel-table gpurun_out/r2r_k_$1_$v.json > gpurun_out/r2r_b_$1_$v.json 2> gpurun_out/r2r_b_$1_$v.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2r_k_$1_$v.json"))
    b = json.loads([l for l in open("gpurun_out/r2r_b_$1_$v.json") if l.startswith("{")][-1])
    print("$1 tiles=$v", [(k["kernel"][:8], round(k["ms"], 4)) for k in d["kernels"]], "ms/step", round(b["ms_per_step"], 4), "dp", b["cpu_baseline"]["max_abs_dp_vs_oracle"])
except Exception as e:
    print("$1 tiles=$v ERR", e)
PY
  done
done
