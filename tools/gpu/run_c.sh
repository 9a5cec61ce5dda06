#!/bin/bash
# round 2, call C: ncu --set full of the rows kernel on the ball
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rows_kernel -s 3 -c 1 -o gpurun_out/r2c_rows \
  python bench.py --workload ball --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/
