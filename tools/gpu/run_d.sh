#!/bin/bash
# round 2, call D: rows kernel v2
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -4 gpurun_out/r2d_pytest.log
for w in ball disk bp2004 eage; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/r2d_kernels_$w.json > gpurun_out/r2d_bench_$w.json 2> gpurun_out/r2d_bench_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2d_kernels_$w.json"))
    print("$w", [(k["kernel"], round(k["ms"], 4)) for k in d["kernels"]])
    b = json.load(open("gpurun_out/r2d_bench_$w.json"))
    print("   ms/step", b["ms_per_step"])
except Exception as e:
    print("$w", "ERR", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rows_kernel -s 3 -c 1 -o gpurun_out/r2d_rows \
  python bench.py --workload ball --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu.log 2>&1
