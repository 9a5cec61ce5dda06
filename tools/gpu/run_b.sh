#!/bin/bash
# round 2, call B: rows kernel (thread per vertex) vs the round-1 lane-group kernel
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
for w in ball disk bp2004 eage; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/r2b_kernels_$w.json > gpurun_out/r2b_bench_$w.json 2> gpurun_out/r2b_bench_$w.err
  DM_ROWS=0 timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/r2b_kernels_${w}_old.json > gpurun_out/r2b_bench_${w}_old.json 2> gpurun_out/r2b_bench_${w}_old.err
  python - <<PY
import json
for tag in ("", "_old"):
    try:
        d = json.load(open("gpurun_out/r2b_kernels_$w%s.json" % tag))
        print("$w", tag or "_new", [(k["kernel"], round(k["ms"], 4)) for k in d["kernels"]])
        b = json.load(open("gpurun_out/r2b_bench_$w%s.json" % tag))
        print("   ms/step", b["ms_per_step"], "maxabs_dp", (b.get("cpu_baseline") or {}).get("max_abs_dp_vs_oracle"))
    except Exception as e:
        print("$w", tag, "ERR", e)
PY
done
