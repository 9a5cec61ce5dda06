#!/bin/bash
# round 2, call E: suite after the pinned-loop / n_rows / relative-displacement changes + ball bench
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -25 gpurun_out/r2e_pytest.log
timeout 300 python bench.py --workload ball --steps 20 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/r2e_kernels_ball.json > gpurun_out/r2e_bench_ball.json 2> gpurun_out/r2e_bench_ball.err
tail -8 gpurun_out/r2e_bench_ball.err
