#!/bin/bash
# vertex_update variants on the gridded workloads: split loops + (A) second interpolation, (B) owner's h through a row search
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
for w in "eage 75 4" "bp2004 25 6" "ball 0.02 0"; do
  set -- $w
  for v in default mirror; do
    if [ $v = mirror ]; then export DM_LIB_PATH=$PWD/build/libdm_mirror.so; else unset DM_LIB_PATH; fi
    F=""; if [ "$3" != "0" ]; then F="--freq $3"; fi
    timeout 300 python bench.py --workload $1 --h0 $2 $F --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel-table gpurun_out/r2k_kernels_$1_$v.json > gpurun_out/r2k_bench_$1_$v.json 2> gpurun_out/r2k_bench_$1_$v.err
    python - <<PY
import json
d = json.load(open("gpurun_out/r2k_kernels_$1_$v.json"))
print("$1", "$v", [(k["kernel"], round(k["ms"], 4)) for k in d["kernels"]])
PY
  done
done
