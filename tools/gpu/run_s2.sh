#!/bin/bash
# ncu --set full of chosen kernels on a chosen workload:  run_s2.sh <tag> <workload> <h0> <freq> <kernel regex> [skip] [count]
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=$1; W=$2; H=$3; FQ=$4; RX=$5; SKIP=${6:-6}; CNT=${7:-2}
F=""; if [ "$FQ" != "0" ]; then F="--freq $FQ"; fi
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT \
  -o gpurun_out/${TAG}_full python bench.py --workload $W --h0 $H $F --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -c 600 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/ | grep ${TAG}
