#!/bin/bash
# GPU suite on both paths + kernel tables of all workloads on both paths
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2y}
for v in 1 0; do
  DM_TILES=$v timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_tiles$v.log 2>&1; echo "tiles=$v pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_tiles$v.log
done
VARIANTS="default buckets" WORKLOADS="${WORKLOADS:-ball:0.02:0 eage:75:4 eage:150:2 bp2004:25:6 bp2004:75:2 disk:0.01:0}" TAG=$TAG bash tools/gpu/run_t.sh
