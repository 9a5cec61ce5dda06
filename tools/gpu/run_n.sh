#!/bin/bash
# launch-configuration sweep on the ball (kernel tables)
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
for v in default cs256 cs64 csmb6 csmb10 vu64 vu256 vumb6 vumb10 ab4 ab6; do
  if [ $v = default ]; then unset DM_LIB_PATH; else export DM_LIB_PATH=$PWD/build/libdm_$v.so; fi
  timeout 300 python bench.py --workload ball --steps 10 --warmup 3 --no-extras --no-cpu-baseline --kernel-table gpurun_out/r2n_kernels_$v.json > gpurun_out/r2n_bench_$v.json 2> gpurun_out/r2n_bench_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2n_kernels_$v.json"))
b = json.loads([l for l in open("gpurun_out/r2n_bench_$v.json") if l.startswith("{")][-1])
print("$v", [(k["kernel"][:8], round(k["ms"], 4)) for k in d["kernels"]], "ms/step", round(b["ms_per_step"], 4))
PY
done
