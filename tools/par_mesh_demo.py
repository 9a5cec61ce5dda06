#!/usr/bin/env python
"""Developer utility: generate_mesh over the ranks of a torchrun launch (NCCL, one GPU per rank).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/par_mesh_demo.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seismicmesh_b200 as sm  # noqa: E402
from seismicmesh_b200 import meshutil  # noqa: E402
from seismicmesh_b200.parallel import TorchComm  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = TorchComm()
arg = sys.argv[1] if len(sys.argv) > 1 else "0.06"
if arg == "eage":  # BASELINE.json configs[4]: the EAGE-shaped sizing grid, slab-decomposed (grid replicated per GPU)
    from bench import synth_vp

    vp, bbox = synth_vp("eage")
    edge = sm.get_sizing_function_from_segy(None, bbox, velocity_data=vp, hmin=150.0, wl=5, freq=2.0, dt=0.001,
                                            grade=0.15, hmax=5e3, domain_pad=250.0, pad_style="linear_ramp",
                                            nz=vp.shape[0], nx=vp.shape[1], ny=vp.shape[2])
    del vp
    dom = sm.Cube(edge.bbox)
else:
    edge = float(arg)
    dom = sm.Cube((0.0, 1.0, 0.0, 2.0, 0.0, 1.0))
t0 = time.perf_counter()
out = sm.generate_mesh(dom, edge, comm=comm, max_iter=25, verbose=0)
dt = time.perf_counter() - t0
if comm.rank == 0:
    p, t = out
    q = meshutil.simp_qual(p, t)
    print(f"ranks={comm.size} N={len(p)} T={len(t)} volume={meshutil.simp_vol(p, t).sum():.4f} "
          f"q mean={q.mean():.4f} wall={dt:.2f}s stats={ {k: (round(v, 3) if isinstance(v, float) else v) for k, v in sm.last_run_stats.items()} }")
else:
    assert out == (True, True)
dist.destroy_process_group()
