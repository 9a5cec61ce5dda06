#!/bin/bash
# build/libdm_<name>.so for every "name=-Dflags" argument (bench them with DM_LIB_PATH=build/libdm_<name>.so)
cd "$(dirname "$0")/.."
for spec in "$@"; do
  name="${spec%%=*}"; defs="${spec#*=}"
  DM_DEFS="$defs" DM_OUT="$PWD/build/libdm_$name.so" bash seismicmesh_b200/csrc/build.sh > /dev/null 2>&1 &
done
wait
ls -la build/
