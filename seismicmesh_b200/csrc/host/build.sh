#!/usr/bin/env bash
# Build libdistmesh_host.so in-tree: the host Delaunay triangulators (2-D, 3-D) behind include/distmesh_host.h.
# -ffp-contract=off: the exact predicates rely on error-free transformations that must not be contracted.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../../libdistmesh_host.so"
"${CXX:-g++}" -O2 -std=c++17 -ffp-contract=off -fPIC -shared -pthread -Wall -I"$HERE/../../../include" \
  "$HERE/dm_delaunay2d.cpp" "$HERE/dm_delaunay3d.cpp" "$HERE/dm_rows.cpp" -o "$OUT"
echo "built $OUT"
