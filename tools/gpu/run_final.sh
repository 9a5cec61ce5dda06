#!/bin/bash
# final check of HEAD: GPU suite with the default layout choice and with each layout forced, then the evidence pack
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2fin}
for v in 1 0; do
  DM_TILES=$v timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_tiles$v.log 2>&1; echo "tiles=$v pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_tiles$v.log; tail -2 gpurun_out/${TAG}_pytest_tiles$v.log
done
TAG=$TAG bash tools/gpu/run_y.sh
