#!/bin/bash
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
for w in "ball 0.02 0" "eage 150 2" "disk 0.01 0"; do
  set -- $w
  F=""; if [ "$3" != "0" ]; then F="--freq $3"; fi
  for v in default abbar default abbar; do
    if [ $v = default ]; then unset DM_LIB_PATH; else export DM_LIB_PATH=$PWD/build/libdm_$v.so; fi
    timeout 300 python bench.py --workload $1 --h0 $2 $F --steps 20 --warmup 3 --no-extras --no-cpu-baseline --kernel-table gpurun_out/r2q_k.json > gpurun_out/r2q_b.json 2> gpurun_out/r2q_b.err
    python - <<PY
import json
d = json.load(open("gpurun_out/r2q_k.json"))
b = json.loads([l for l in open("gpurun_out/r2q_b.json") if l.startswith("{")][-1])
print("$1 $v", [(k["kernel"][:8], round(k["ms"], 4)) for k in d["kernels"]], "ms/step", round(b["ms_per_step"], 4))
PY
  done
done
