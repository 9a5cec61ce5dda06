#!/bin/bash
# executed CALL instructions (special-case subroutines of fp64 sqrt / division) per kernel of one step
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
W=$1; H=$2; FQ=$3; TAG=$4
F=""; if [ "$FQ" != "0" ]; then F="--freq $FQ"; fi
timeout 900 ncu --section SourceCounters --import-source on -k regex:'prep_kernel|cull_scatter_kernel|cull_bin_kernel|adjacency_kernel|tile_rows_kernel|vertex_update_kernel|project_list_kernel' -s 20 -c 5 -o gpurun_out/${TAG} python bench.py --workload $W --h0 $H $F --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}.log 2>&1
ls -la gpurun_out/${TAG}.ncu-rep
