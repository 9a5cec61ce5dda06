#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python tools/gpu/ttol_sweep.py > gpurun_out/r2g_ttol.log 2>&1
tail -40 gpurun_out/r2g_ttol.log
