#!/bin/bash
# A/B: Newton projection fused into vertex_update (build/libdm_fusep.so) against the default library
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2cq}
for w in "ball 0.02 0" "disk 0.01 0" "eage 150 2" "bp2004 75 2" "eage 75 4"; do
  set -- $w
  F=""; if [ "$3" != "0" ]; then F="--freq $3"; fi
  for v in default fusep default fusep; do
    if [ $v = default ]; then unset DM_LIB_PATH; else export DM_LIB_PATH=$PWD/build/libdm_$v.so; fi
    timeout 300 python bench.py --workload $1 --h0 $2 $F --steps 20 --warmup 3 --no-extras --kernel-table gpurun_out/${TAG}_k.json > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_b.err
    python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_k.json"))
b = json.loads([l for l in open("gpurun_out/${TAG}_b.json") if l.startswith("{")][-1])
print("$1 $2 $v", [(k["kernel"][:8], round(k["ms"], 4)) for k in d["kernels"]], "ms/step", round(b["ms_per_step"], 4), "dp", b["cpu_baseline"]["max_abs_dp_vs_oracle"])
PY
  done
done
export DM_LIB_PATH=$PWD/build/libdm_fusep.so
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
