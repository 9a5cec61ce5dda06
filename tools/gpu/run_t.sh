#!/bin/bash
# variants of the tile layout: kernel tables on ball / eage150 (+ smoke of the default build)
cd "$GRAFT_REPO_ROOT"
export DM_BENCH_CACHE=/tmp/dmcache
TAG=${TAG:-r2t}
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
WL=${WORKLOADS:-ball:0.02:0 eage:150:2}
for w in $WL; do
  IFS=: read W H FQ <<< "$w"
  F=""; if [ "$FQ" != "0" ]; then F="--freq $FQ"; fi
  for v in ${VARIANTS:-default}; do
    if [ $v = default ]; then unset DM_LIB_PATH; unset DM_TILES; elif [ $v = buckets ]; then unset DM_LIB_PATH; export DM_TILES=0; else unset DM_TILES; export DM_LIB_PATH=$PWD/build/libdm_$v.so; fi
    timeout 300 python bench.py --workload $W --h0 $H $F --steps 20 --warmup 3 --no-extras --no-cpu-baseline --kernel-table gpurun_out/${TAG}_k.json > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_b.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_k.json"))
    b = json.loads([l for l in open("gpurun_out/${TAG}_b.json") if l.startswith("{")][-1])
    print("$W $v", [(k["kernel"][:8], round(k["ms"], 4)) for k in d["kernels"]], "ms/step", round(b["ms_per_step"], 4))
except Exception as e:
    print("$W $v ERR", e)
PY
  done
done
