#!/bin/bash
# box-size sweep of the threaded 3-D triangulator on the box's cores
cd "$GRAFT_REPO_ROOT"
TAG=${TAG:-r2cc}
python - <<'PY'
import numpy as np, sys
sys.path.insert(0,'.')
from seismicmesh_b200.generation import _staggered_grid
for h0,name in ((0.05,'ball005'),(0.035,'ball0035'),(0.02,'ball002')):
    s=np.ascontiguousarray(_staggered_grid(h0,3,np.array([[-1.,1.]]*3)))
    p=s[(s**2).sum(1)<1.0]
    p=p+np.random.default_rng(0).uniform(-0.1*h0,0.1*h0,p.shape)
    p.tofile('/tmp/%s.bin'%name)
PY
g++ -O2 -std=c++17 -ffp-contract=off -pthread -Iinclude -Iseismicmesh_b200/csrc/host tools/host/delaunay3d_file_driver.cpp seismicmesh_b200/csrc/host/dm_delaunay3d.cpp -o /tmp/drvf
for f in ball005 ball0035 ball002; do for th in 1 8 12 15; do for rr in ${PASSES:-5 9}; do for pr in 1000 1500 2500; do
  echo "== $f threads $th passes $rr pass_rows $pr: $(DM_HOST_PASSES=$rr DM_HOST_PASS_ROWS=$pr /tmp/drvf /tmp/$f.bin $th 2>&1 | tail -1)"
  [ $th = 1 ] && break 2
done; done; done; done > gpurun_out/${TAG}_tune.log 2>&1
