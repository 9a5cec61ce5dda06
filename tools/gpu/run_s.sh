#!/bin/bash
# ncu --set full of the tile-layout kernels on the ball
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${1:-r2s}
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'cull_bin_kernel|tile_rows_kernel' -s 6 -c 2 \
  -o gpurun_out/${TAG}_full python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/ | grep ${TAG}
