#!/usr/bin/env bash
# Build libdistmesh_b200.so in-tree for sm_100a (travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${DM_OUT:-$HERE/../libdistmesh_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
  -Xcompiler -fPIC,-O2,-Wall -shared ${DM_PTXAS_V:+-Xptxas -v} ${DM_DEFS:-} \
  -I"$HERE/../../include" "$HERE/dm_kernels.cu" -o "$OUT"
echo "built $OUT"
