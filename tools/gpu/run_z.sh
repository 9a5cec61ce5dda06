#!/bin/bash
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2af}
timeout 900 python -m pytest tests -m gpu -x -q -k "${K:-size_eval or force_iteration or hub or layout or reuse or gridded or sizing}" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
VARIANTS="${VARIANTS:-default}" WORKLOADS="${WORKLOADS:-eage:75:4 eage:150:2 bp2004:25:6 bp2004:75:2}" TAG=$TAG bash tools/gpu/run_t.sh
