#!/bin/bash
# round 2, call A: full GPU suite at HEAD + baseline benches + sanitizer on smoke
cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_ball.json 2> gpurun_out/r2a_bench_ball.err; tail -c 600 gpurun_out/r2a_bench_ball.json
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_memcheck.log 2>&1; tail -3 gpurun_out/r2a_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_racecheck.log 2>&1; tail -3 gpurun_out/r2a_racecheck.log
