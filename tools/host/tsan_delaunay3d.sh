#!/usr/bin/env bash
# ThreadSanitizer over the threaded 3-D triangulator (dmh_delaunay3d_mt): builds the driver with
# -fsanitize=thread and runs it on N random points in a ball with THREADS threads (three calls, cell
# list compared with the serial call).  Usage: tools/host/tsan_delaunay3d.sh [N] [THREADS] > log
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$HERE/../.."
OUT="$(mktemp -d)"
g++ -O1 -g -fsanitize=thread -std=c++17 -ffp-contract=off -pthread -I"$ROOT/include" -I"$ROOT/seismicmesh_b200/csrc/host" \
  "$HERE/delaunay3d_driver.cpp" "$ROOT/seismicmesh_b200/csrc/host/dm_delaunay3d.cpp" -o "$OUT/drv_tsan"
echo "ThreadSanitizer: dmh_delaunay3d_mt, N=${1:-120000}, threads=${2:-8}"
"$OUT/drv_tsan" "${1:-120000}" "${2:-8}" 2>&1
echo "exit code $? (ThreadSanitizer prints 'WARNING: ThreadSanitizer: data race' and exits 66 on a report)"
