#!/bin/bash
cd "$GRAFT_REPO_ROOT"
bash tools/gpu/profile.sh r2p
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_memcheck.log 2>&1; tail -3 gpurun_out/r2p_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_racecheck.log 2>&1; tail -3 gpurun_out/r2p_racecheck.log
timeout 600 compute-sanitizer --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_synccheck.log 2>&1; tail -3 gpurun_out/r2p_synccheck.log
