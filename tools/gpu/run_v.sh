#!/bin/bash
# GPU suite (default layouts) + default bench line + per-workload kernel tables
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2aa}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
VARIANTS="default" WORKLOADS="ball:0.02:0 eage:75:4 bp2004:25:6" TAG=$TAG bash tools/gpu/run_t.sh
