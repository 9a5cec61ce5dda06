#!/bin/bash
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
timeout 900 python -m pytest tests -m gpu -x -q -k "grid or reuse or hub or staged or segy or force_iteration" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -4 gpurun_out/r2m_pytest.log
for w in "eage 75 4" "bp2004 25 6" "eage 150 2" "bp2004 75 2"; do
  set -- $w
  for v in mirror hall; do
    if [ $v = hall ]; then export DM_LIB_PATH=$PWD/build/libdm_hall.so; else unset DM_LIB_PATH; fi
    timeout 300 python bench.py --workload $1 --h0 $2 --freq $3 --steps 10 --warmup 3 --no-extras --kernel-table gpurun_out/r2m_kernels_$1_$2_$v.json > gpurun_out/r2m_bench_$1_$2_$v.json 2> gpurun_out/r2m_bench_$1_$2_$v.err
    python - <<PY
import json
d = json.load(open("gpurun_out/r2m_kernels_$1_$2_$v.json"))
b = json.loads([l for l in open("gpurun_out/r2m_bench_$1_$2_$v.json") if l.startswith("{")][-1])
print("$1 $2 $v", [(k["kernel"], round(k["ms"], 4)) for k in d["kernels"]], "ms/step", round(b["ms_per_step"], 4), "dp", b["cpu_baseline"]["max_abs_dp_vs_oracle"])
PY
  done
done
