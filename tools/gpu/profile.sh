#!/bin/bash
# ncu evidence of the CURRENT build (run under gpurun; outputs land in gpurun_out/, copy the summaries to
# profiles/ with tools/make_ncu_traffic.py):
#   1. launch list of the default bench command (shares of the step, cold-cache and serialised);
#   2. one `--set full` capture of every kernel of ONE force iteration on the headline ball workload.
# usage: bash tools/gpu/profile.sh <tag>
cd "$GRAFT_REPO_ROOT"
TAG=${1:-r2}
export DM_BENCH_CACHE=/tmp/dmcache
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
# the full-set capture: skip the set-up + warm-up launches, take the five kernels of one timed step
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'prep_kernel|cull_scatter_kernel|cull_bin_kernel|adjacency_kernel|tile_rows_kernel|vertex_update_kernel|project_list_kernel' -s 20 -c 5 \
  -o gpurun_out/${TAG}_full python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
