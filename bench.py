#!/usr/bin/env python
"""bench.py -- DistMesh force-iteration throughput (vertex-updates/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ball|disk|bp2004|eage] [--impl ours|reference]

A *step* is one pass of the hot path (cull -> unique bars -> forces -> update -> projection ->
max|dp|, i.e. one iteration of the loop body of SeismicMesh generate_mesh AFTER its Delaunay
call) over one synthetic (p, t).  Delaunay is host work in the reference's design and is done
once in the untimed set-up; it is reported separately (`delaunay_s`).

Prints ONE JSON line (contract in the task statement): `value` = vertex-updates/s with inputs
resident in HBM, `e2e` = the same through the public host-buffer call generate_mesh itself uses
(ForceLoop.iterate_host: H2D of p and t from pinned memory + D2H of the new positions inside the timed
region), `roofline` for the dominant kernel (algorithmic bytes / CUDA-event time against
MEASURED_PEAKS.json), `cpu_baseline` = the oracle port of the reference's NumPy loop body on one host core.

The default invocation (N = 1) also carries, as sub-records of the same line:
  `workloads`     the north_star target workloads: BP2004-shaped (hmin 25 m, 6 Hz) and EAGE-shaped
                  (hmin 75 m, 4 Hz), gridded fh -- value, e2e, roofline, max|dp| against the oracle;
  `time_to_mesh`  BASELINE.json's second metric: generate_mesh end to end (host Delaunay included) on
                  disk h0 = 0.01 and ball h0 = 0.05, 25 iterations, reference semantics and `ttol`.
With N > 1 the step is the slab-decomposed one (one slab per GPU, ghosts chosen by dm_halo_select, one
halo exchange per step), on the ball-sized cylinder slabs and, as a sub-record, on EAGE-shaped slabs; the
`time_to_mesh` block is then generate_mesh(comm=...) over the N ranks (a box of N unit cubes at
h0 = 0.031, and the EAGE-shaped hmin 150 m domain cut into N slabs).
`--impl reference` times the reference's own CPU loop body (oracle port + the reference's compiled
unique_edges) on the SAME configuration, without importing the product package.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DistMesh vertex-updates/sec (verts x iters / s), device-resident force iteration"
UNIT = "vertex-updates/s"
TRIANGULATOR = None  # --triangulator: None = the package default (native), "qhull", "native"


# ------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs) -- the builders in this block use NumPy / SciPy
# only, so the reference arm can share them without importing the product package
# ------------------------------------------------------------------------------------------------
def _lattice(h0, dim, bbox):
    axes = [np.arange(int(np.ceil((hi + h0 - lo) / h0)), dtype=float) * h0 + lo for lo, hi in bbox]
    g = [a.copy() for a in np.meshgrid(*axes, indexing="ij")]
    g[1][1::2] += h0 / 2
    if dim == 3:
        g[2][1::2] += h0 / 2
    return np.stack([a.ravel() for a in g], axis=1)


def make_points(workload, h0, seed=0):
    """Synthetic input of BASELINE.json configs[1] (ball) / configs[0] (disk): the reference's
    initial lattice inside the domain (uniform h => no rejection), jittered by 0.1*h0 (seeded) so
    the host Delaunay is non-degenerate, as in a mid-run iteration."""
    dim = 3 if workload == "ball" else 2
    bbox = np.array([[-1.0, 1.0]] * dim)
    p = _lattice(h0, dim, bbox)
    r = np.sqrt((p**2).sum(1))
    p = p[r - 1.0 < 0.1 * h0]
    rng = np.random.default_rng(seed)
    p = p + rng.uniform(-0.1 * h0, 0.1 * h0, p.shape)
    return np.ascontiguousarray(p), dim


def make_cylinder_points(workload, h0, world, seed=0):
    """Weak-scaling input for N > 1: `world` slabs along axis 1 of a cylinder (3-D) / rectangle (2-D),
    each slab with the volume (area) of the unit ball (disk), i.e. the per-GPU work of the N = 1 workload."""
    if workload == "ball":
        dim, ell = 3, 4.0 / 3.0
    else:
        dim, ell = 2, np.pi / 2.0
    Ly = world * ell
    lo, hi = [-1.0] * dim, [1.0] * dim
    lo[1], hi[1] = -Ly / 2, Ly / 2
    p = _lattice(h0, dim, list(zip(lo, hi)))
    if dim == 3:
        rad = np.sqrt(p[:, 0] ** 2 + p[:, 2] ** 2)
        inside = (rad - 1.0 < 0.1 * h0) & (np.abs(p[:, 1]) - Ly / 2 < 0.1 * h0)
    else:
        inside = (np.abs(p[:, 0]) - 1.0 < 0.1 * h0) & (np.abs(p[:, 1]) - Ly / 2 < 0.1 * h0)
    p = p[inside]
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(p + rng.uniform(-0.1 * h0, 0.1 * h0, p.shape)), dim, Ly


def synth_vp(workload):
    """Synthetic velocity models of BASELINE.json configs[2] / configs[3]: the
    named grid shapes and physical extents, layered background + water layer + salt body."""
    if workload == "bp2004":
        nz, nx = 1911, 5395
        bbox = (-12000.0, 0.0, 0.0, 67000.0)
        z = np.linspace(bbox[0], bbox[1], nz)[:, None]
        x = np.linspace(bbox[2], bbox[3], nx)[None, :]
        vp = 1500 + (-z / 12000) * 3000 + 150 * np.sin(x / 3000) * np.cos(z / 1500)
        vp = np.where(z > -1000 - 300 * np.sin(x / 8000), 1486.0, vp)
        vp = np.where(((x - 30000) / 9000) ** 2 + ((z + 6000) / 2500) ** 2 < 1, 4790.0, vp)
        return np.ascontiguousarray(vp), bbox
    nz, nx, ny = 210, 676, 676
    bbox = (-4200.0, 0.0, 0.0, 13520.0, 0.0, 13520.0)
    z = np.linspace(bbox[0], bbox[1], nz)[:, None, None]
    x = np.linspace(bbox[2], bbox[3], nx)[None, :, None]
    y = np.linspace(bbox[4], bbox[5], ny)[None, None, :]
    vp = 1500 + (-z / 4200) * 2800 + 120 * np.sin(x / 2500) * np.cos(y / 2000) + 0 * z
    r2 = ((x - 6500) / 3000) ** 2 + ((y - 7000) / 2600) ** 2 + ((z + 2300) / 900) ** 2
    vp = np.where(r2 < 1, 4480.0, vp)
    return np.ascontiguousarray(vp), bbox


def sizing_kwargs(workload, vp, h0=None, freq=None):
    if workload == "bp2004":  # README.md:176-186, benchmarks/benchmark_BP2004.py:31-39
        hmin, fr = h0 or 75.0, freq or 2.0
        return hmin, fr, 2, dict(hmin=hmin, wl=10, freq=fr, dt=0.001, grade=0.15, domain_pad=1e3, pad_style="edge",
                                 nz=vp.shape[0], nx=vp.shape[1])
    hmin, fr = h0 or 150.0, freq or 2.0  # README.md:275-292
    return hmin, fr, 3, dict(hmin=hmin, wl=5, freq=fr, dt=0.001, grade=0.15, hmax=5e3, domain_pad=250.0,
                             pad_style="linear_ramp", nz=vp.shape[0], nx=vp.shape[1], ny=vp.shape[2])


def domain_spec(workload, bbox=None):
    if workload == "ball":
        return ("ball", dict(x0=[0.0, 0.0, 0.0], r=1.0))
    if workload == "disk":
        return ("disk", dict(x0=[0.0, 0.0], r=1.0))
    return ("rectangle", dict(bbox=tuple(bbox))) if len(bbox) == 4 else ("cube", dict(bbox=tuple(bbox)))


# ------------------------------------------------------------------------------------------------
# product-side set-up
# ------------------------------------------------------------------------------------------------
def delaunay_backend(dim):
    from seismicmesh_b200.triangulator import get_triangulator

    return get_triangulator(TRIANGULATOR, dim).name


def triangulate(p):
    """Host Delaunay with the package's default (native) triangulator.  DM_BENCH_CACHE=<dir> re-uses the
    cells of an identical point set between runs of one profiling session (the set-up is untimed)."""
    from seismicmesh_b200.triangulator import get_triangulator

    cache = os.environ.get("DM_BENCH_CACHE")
    key = None
    if cache:
        import hashlib

        key = os.path.join(cache, f"tri2_{TRIANGULATOR or 'default'}_" + hashlib.sha1(p.tobytes()).hexdigest()[:16] + ".npz")
        if os.path.exists(key):
            z = np.load(key)
            return z["t"], float(z["dt"])
    t0 = time.perf_counter()
    t = get_triangulator(TRIANGULATOR, p.shape[1]).triangulate(p)
    dt = time.perf_counter() - t0
    if key:
        os.makedirs(cache, exist_ok=True)
        np.savez(key, t=t, dt=dt)
    return t, dt


def build_workload(workload, h0=None, freq=None, settle=2):
    """-> dict(p, t, dim, dom, size, h0, spec, fh_grid, delaunay_s, desc).  ball / disk: the reference's
    lattice inside the SDF (uniform h).  bp2004 / eage: gridded fh from our own
    get_sizing_function_from_segy on the synthetic velocity model (device gradient limiter), the
    reference's rejection-sampled initial points, then `settle` real iterations so that the timed
    (p, t) is a mid-run state (non-degenerate Delaunay), as for the jittered ball."""
    import seismicmesh_b200 as sm
    from seismicmesh_b200.engine import SizeSpec

    if workload in ("ball", "disk"):
        h0 = h0 or (0.02 if workload == "ball" else 0.01)
        p, dim = make_points(workload, h0)
        dom = sm.Ball([0.0, 0.0, 0.0], 1.0) if dim == 3 else sm.Disk([0.0, 0.0], 1.0)
        t, dt = triangulate(p)
        return dict(p=p, t=t, dim=dim, dom=dom, size=SizeSpec(dim, const=h0), h0=h0, spec=domain_spec(workload), fh_grid=None,
                    delaunay_s=dt, desc=f"{workload}_h0={h0:g}", sizing_s=0.0, edge=h0)
    import torch

    from seismicmesh_b200 import device as D
    from seismicmesh_b200.engine import ForceLoop, Level
    from seismicmesh_b200.generation import _initial_points

    vp, bbox = synth_vp(workload)
    hmin, fr, dim, kw = sizing_kwargs(workload, vp, h0, freq)
    t0 = time.perf_counter()
    ef = sm.get_sizing_function_from_segy(None, bbox, velocity_data=vp, **kw)
    sizing_s = time.perf_counter() - t0
    del vp
    dom = sm.Rectangle(ef.bbox) if dim == 2 else sm.Cube(ef.bbox)
    size = SizeSpec(dim, interp=ef.interpolant())
    geps, deps = 0.1 * hmin, np.sqrt(np.finfo(np.double).eps) * hmin
    level = Level(dom, dim)
    p = _initial_points(hmin, geps, dim, np.array(ef.bbox).reshape(-1, 2), size, level, np.empty((0, dim)),
                        dict(r0m_is_h0=False, seed=0))
    p = np.ascontiguousarray(p)
    loop = ForceLoop(dim, [level], size, hmin, geps, deps)
    dt = 0.0
    for _ in range(settle):
        t, dt = triangulate(p)
        p = loop.iterate(D.to_dev(p, torch.float64), D.to_dev(t, torch.int32))[0].cpu().numpy()
    t, dt = triangulate(p)
    g = ef.interpolant()
    return dict(p=p, t=t, dim=dim, dom=dom, size=size, h0=hmin, spec=domain_spec(workload, ef.bbox), fh_grid=(g.grid, g.values),
                delaunay_s=dt, sizing_s=sizing_s, edge=ef,
                desc=f"{workload}_shaped_grid{'x'.join(str(n) for n in g.values.shape)}_hmin={hmin:g}_freq={fr:g}")


def build_slab_workload(workload, rank, world, h0=None, freq=None):
    """One slab per GPU through the package's own parallel path (generate_mesh(comm=...) stopped before
    its gather): every rank triangulates its owned vertices, dm_halo_select picks the vertices whose
    incident-cell circumballs reach a neighbour's extent, the neighbours receive them as ghosts and
    the slab is retriangulated with them -- exactly what a parallel iteration feeds the device.
    ball / disk: cylinder / rectangle slabs with the N = 1 work per GPU (restart from the jittered
    lattice); eage: the EAGE-shaped domain cut along axis 1 with hmin / freq scaled so that the work per
    GPU stays that of the hmin 75 m / 4 Hz single-GPU case (BASELINE.json configs[4])."""
    import seismicmesh_b200 as sm
    from seismicmesh_b200.engine import SizeSpec
    from seismicmesh_b200.parallel import SlabLayout, TorchComm

    comm = TorchComm()
    sizing_s = 0.0
    if workload in ("ball", "disk"):
        h0 = h0 or (0.02 if workload == "ball" else 0.01)
        pts, dim, Ly = make_cylinder_points(workload, h0, world)
        dom = sm.Cylinder(h=Ly, r=1.0) if dim == 3 else sm.Rectangle((-1.0, 1.0, -Ly / 2, Ly / 2))
        edge, size, fh_grid = h0, SizeSpec(dim, const=h0), None
        # `axis=0` is decomp.blocker's name for cuts along y (dimension 1), the cylinder's axis
        st = sm.generate_mesh(dom, edge, comm=comm, points=pts, max_iter=1, axis=0, verbose=0, _return_state=True)
        desc = f"{workload}_h0={h0:g} x {world} slabs (cylinder of {world} ball volumes)"
    else:
        vp, bbox = synth_vp(workload)
        sc = world ** (1.0 / 3.0)
        hmin, fr, dim, kw = sizing_kwargs(workload, vp, (h0 or 75.0) / sc, (freq or 4.0) * sc)
        t0 = time.perf_counter()
        ef = sm.get_sizing_function_from_segy(None, bbox, velocity_data=vp, **kw)
        sizing_s = time.perf_counter() - t0
        del vp
        dom = sm.Cube(ef.bbox)
        edge, size = ef, SizeSpec(dim, interp=ef.interpolant())
        g = ef.interpolant()
        fh_grid = (g.grid, g.values)
        h0 = hmin
        st = sm.generate_mesh(dom, edge, comm=comm, max_iter=3, axis=1, verbose=0, _return_state=True)
        desc = f"eage_shaped_grid{'x'.join(str(n) for n in g.values.shape)}_hmin={hmin:.4g}_freq={fr:.4g} x {world} slabs along axis 1"
    n_own = st["n_owned"]
    layout = SlabLayout(np.arange(n_own), np.zeros(st["n_ghost_below"], dtype=np.int64), np.zeros(st["n_ghost_above"], dtype=np.int64),
                        st["export_below"].astype(np.int64), st["export_above"].astype(np.int64))
    stats = dict(sm.last_run_stats)
    return dict(p=np.ascontiguousarray(st["p"]), t=np.ascontiguousarray(st["t"], dtype=np.int32), dim=dim, dom=dom, size=size, h0=h0,
                spec=None, fh_grid=fh_grid, delaunay_s=stats.get("delaunay", 0.0), sizing_s=sizing_s, edge=edge, desc=desc,
                layout=layout, n_owned=n_own)


def oracle_step_fn(spec, h0, fh_grid):
    """The reference's loop body (oracle port, + the reference's own native unique_edges from
    oracle/_ref when it was built) as a closure step(p, t) -> p_new."""
    from oracle import distmesh_oracle as orc
    from oracle import ref_harness

    native = None
    if ref_harness.native_available():
        try:
            native = ref_harness.load_native()
        except Exception:
            native = None
    if native is not None:
        def unique_bars(t):  # mesh_generator.py:680-688 calling the reference's compiled unique_edges
            e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
            if t.shape[1] == 4:
                e = np.concatenate((e, t[:, [0, 3]], t[:, [1, 3]], t[:, [2, 3]]), axis=0)
            return native.unique_edges(e)
        orc.unique_bars = unique_bars
    geps, deps = 0.1 * h0, np.sqrt(np.finfo(np.double).eps) * h0
    fd = lambda x: orc.sdf(spec, x)  # noqa: E731
    if fh_grid is None:
        fh = lambda x: np.array([h0] * len(x))  # noqa: E731
    else:
        axes, grid = fh_grid
        fh = lambda x: orc.interp_grid(list(axes), grid, x)  # noqa: E731

    def step(p, t):
        return orc.force_iteration(p, t, [fd], fh, h0, geps, deps)["p"]

    return step, ("reference-native unique_edges + NumPy port" if native is not None else "NumPy port")


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(N, T, Tk, E, dim, grid_fh=False):
    """Compulsory HBM traffic per kernel launch (each input read once, each output written once,
    gathers counted once per distinct element; DESIGN.md section 4 / SURVEY section 8d), from the
    ACTUAL sizes of the run."""
    c, d = dim + 1, dim
    hs = 16 * E if grid_fh else 0  # gridded fh: h is stored at both slots of a bar (8 B each), written once, read once
    return {
        # zero the per-iteration counters; 3-D: p read once, the padded copy written once
        "prep(zero+pad)": 4 * (N + 1) + (56 * N if dim == 3 else 0),
        # t and p read once, keep flags + the kept cells handed to the bar stage written once
        "cull_scatter": 4 * c * T + 8 * d * N + T + 4 * c * Tk,
        # SURVEY 8d K2 (+K5: rows are symmetric, every bar is stored at both of its ends) with the
        # bar pass K3 fused in (positions read once, h written once per bar for gridded fh)
        "adjacency": 4 * c * Tk + 8 * (N + 1) + 8 * E + 8 * d * N + hs,
        # the tile layout of stages A + B (dm_tiles.cuh): the same compulsory traffic
        "cull_bin": 4 * c * T + 8 * d * N + T + 4 * c * Tk,
        "tile_rows": 4 * c * Tk + 8 * (N + 1) + 8 * E + 8 * d * N + hs,
        "bar_pass+scale": 8 * d * N + 4 * E + 8 * N + hs,
        "vertex_update+maxdp": 16 * d * N + 8 * E + 8 * N + hs,
    }


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed
    `ncu --set full` capture of THIS build on this workload (profiles/ncu_traffic.json, regenerated by
    tools/gpu/profile.sh; the file names the capture it came from); None if there is none."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload, {}).get(kernel)
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# the reference arm: the reference's own CPU loop body on the SAME configuration, no product imports
# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    from scipy.spatial import Delaunay

    K, W = args.steps, args.warmup
    workload = args.workload
    if workload not in ("ball", "disk"):
        print(json.dumps({"impl": "reference", "unavailable":
                          "gridded workloads need the sizing preprocessing; the reference arm runs the headline ball / disk configurations"}))
        return
    h0 = args.h0 or (0.02 if workload == "ball" else 0.01)
    p, dim = make_points(workload, h0)  # the GPU arm's exact point set (same seed, same jitter)
    t0 = time.perf_counter()
    t = np.ascontiguousarray(Delaunay(p).simplices, dtype=np.int32)  # Qhull stands in for the reference's CGAL
    tq = time.perf_counter() - t0
    step, kind = oracle_step_fn(domain_spec(workload), h0, None)
    # a full-size ball step is ~3 s on one core: bound the run to ~4 minutes, say how many steps were taken
    c0 = time.perf_counter()
    step(p, t)
    one = time.perf_counter() - c0
    Wd = max(0, min(W, int(30.0 // max(one, 1e-9))) - 1)
    Kd = max(1, min(K, int(200.0 // max(one, 1e-9))))
    for _ in range(Wd):
        step(p, t)
    c0 = time.perf_counter()
    for _ in range(Kd):
        step(p, t)
    dt = time.perf_counter() - c0
    N = len(p)
    val = N * Kd / dt
    sample = f"{workload}_h0={h0:g} (N={N}, T={len(t)}): {Kd} timed steps of the full loop body after {Wd + 1} warm-up steps, 1 core"
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": Kd, "warmup": Wd + 1,
        "ms_per_step": dt / Kd * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{workload}_h0={h0:g}", "N_per_gpu": N, "T_per_gpu": int(len(t)), "dim": dim, "same_config": True,
                   "delaunay": "excluded (host, set-up; Qhull here, CGAL in the reference)",
                   "note": "kind=port: the NumPy restatement of the reference's loop body with the reference's own compiled "
                           "unique_edges; the unmodified reference measured through its public API is ~2.3x slower "
                           "(BASELINE.md section 2), so this baseline is the conservative one"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "detail": kind},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "delaunay_s": tq,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def timed(fn, K, flush, barrier):
    import torch

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for i in range(K):
        flush.fill_(i & 0xFF)  # L2 flush between timed iterations (not timed)
        ev[i][0].record()
        fn()
        ev[i][1].record()
    barrier()
    return float(sum(a.elapsed_time(b) for a, b in ev))


def measure(wl, K, W, rank, world, local_rank, full, flush, kernel_table=None, cpu_check=True):
    """The timed legs on one workload.  full: also the clocks sampler, the opt-in row-reuse step, the
    sliver pass and the timed CPU baseline (the headline record); otherwise the sub-record subset."""
    import torch
    import torch.distributed as dist

    from seismicmesh_b200 import device as D
    from seismicmesh_b200._lib import check, lib
    from seismicmesh_b200.engine import ForceLoop, Level

    p, t, dim, dom, h0, size = wl["p"], wl["t"], wl["dim"], wl["dom"], wl["h0"], wl["size"]
    layout = wl.get("layout")
    n_owned = wl.get("n_owned", len(p))
    N, T = len(p), len(t)
    geps, deps = 0.1 * h0, np.sqrt(np.finfo(np.double).eps) * h0
    loop = ForceLoop(dim, [Level(dom, dim)], size, h0, geps, deps)
    if layout is not None:
        loop.n_rows = n_owned
    p_dev = D.to_dev(p, torch.float64)
    t_dev = D.to_dev(t, torch.int32)
    p_out = p_dev.clone()  # (ghost rows of the output are only ever written by the halo exchange)

    halo, halo_kind, direct = None, None, None
    if layout is not None:
        from seismicmesh_b200.parallel import DirectHalo, PeerHalo, RingHalo

        ring = RingHalo(layout, dim, p_dev.device, rank=rank, world=world)
        halo, halo_kind = ring, "NCCL P2P send/recv (RingHalo)"
        mode = os.environ.get("DM_HALO", "direct")
        if mode in ("direct", "p2p"):
            try:  # NVLink peer-memory exchange with our own kernels; checked once against the NCCL exchange
                a = p_dev.clone()
                a[n_owned:] = float("nan")
                ring.exchange(a)
                if mode == "direct":
                    direct = DirectHalo(layout, dim, p_dev.device, rank=rank, world=world)
                    for _ in range(2):  # both slots
                        b = direct.slot()
                        b.copy_(p_dev)
                        b[n_owned:] = float("nan")
                        torch.cuda.synchronize()
                        dist.barrier()
                        direct.exchange(b)
                        torch.cuda.synchronize()
                        direct.check()
                        assert torch.equal(a, b), "DirectHalo and RingHalo disagree"
                    halo, halo_kind = direct, ("NVLink stores straight into the neighbours' ghost rows + stamps, 2 launches "
                                               "(dm_halo_push2 + dm_halo_wait; DirectHalo)")
                else:
                    peer = PeerHalo(layout, dim, p_dev.device, rank=rank, world=world)
                    b = p_dev.clone()
                    b[n_owned:] = float("nan")
                    peer.exchange(b)
                    peer.exchange(b)  # both slots
                    torch.cuda.synchronize()
                    assert torch.equal(a, b), "PeerHalo and RingHalo disagree"
                    halo, halo_kind = peer, "NVLink peer-memory push, dm_halo_push + symmetric-memory signals (PeerHalo)"
            except Exception as exc:  # no peer access / symmetric memory on this box: keep NCCL
                direct = None
                halo_kind += f" [peer-memory path unavailable: {type(exc).__name__}: {exc}]"

    def one_step():
        if direct is not None:  # the result goes to this step's symmetric slot, whose ghost rows the neighbours fill
            out = direct.slot()
            loop.iterate(p_dev, t_dev, p_out=out)
            direct.exchange(out)
            return
        loop.iterate(p_dev, t_dev, p_out=p_out)
        if halo is not None:
            halo.exchange(p_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=p_dev.device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # pinned host buffers of the e2e leg (allocated before the clocks sampler starts)
    p_pin = torch.from_numpy(p).pin_memory()
    t_pin = torch.from_numpy(t).pin_memory()
    out_pin = torch.empty_like(p_pin).pin_memory()
    sc_pin = torch.empty(8, dtype=torch.float64).pin_memory()
    for _ in range(W):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank) if (full and rank == 0) else None
    if sampler:
        sampler.start()
    wall0 = time.perf_counter()
    total_ms = allmax(timed(one_step, K, flush, barrier))
    wall = time.perf_counter() - wall0
    if world > 1:
        nn = torch.tensor([n_owned], dtype=torch.float64, device=p_dev.device)
        dist.all_reduce(nn, op=dist.ReduceOp.SUM)
        N_all = int(nn.item())
    else:
        N_all = n_owned
    value = N_all * K / (total_ms * 1e-3)

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region: the call
    #      generate_mesh makes every iteration (there with the positions already resident)
    def e2e_step():
        q = loop.iterate_host(p_pin, t_pin, out_pin)
        if halo is not None:
            halo.exchange(q)
        sc_pin.copy_(loop.plan.scalars(), non_blocking=True)

    for _ in range(2):
        e2e_step()
    e2e_ms = allmax(timed(e2e_step, K, flush, barrier))
    e2e_value = N_all * K / (e2e_ms * 1e-3)
    clocks = sampler.stop() if sampler else None  # sampled every 20 ms across both timed regions
    maxdp = float(sc_pin[4])
    h2d = int(p.nbytes + t.nbytes)
    rec = {
        "value": value, "unit": UNIT, "ms_per_step": total_ms / K,
        "config": {"workload": wl["desc"], "N_per_gpu": N, "T_per_gpu": T, "dim": dim,
                   "l2": "flushed between timed steps (512 MiB fill, untimed)", "delaunay": "host, set-up (untimed)",
                   "parallelism": (f"{world} slabs (owned + dm_halo_select ghosts per GPU; rows / forces for owned vertices only), "
                                   f"halo exchange per step: {halo_kind}; halo bytes/step/rank={halo.bytes_per_exchange}; "
                                   f"owned={n_owned} ghosts={N - n_owned} on rank 0") if layout is not None else "single",
                   "N_owned_total": N_all,
                   "stage_ab_layout": f"{loop.layout} (include/distmesh_b200.h DM_LAYOUT_*; the kernel table names the kernels that ran)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(p.nbytes + 64),
                "ms_per_step": e2e_ms / K, "h2d_gbs_per_rank": h2d / (e2e_ms / K * 1e-3) / 1e9,
                "call": "ForceLoop.iterate_host (the call generate_mesh makes every iteration): pinned H2D of p and t, t in "
                        "chunks overlapped with stage A, async D2H of the new positions"},
        "gpu_launches": 5 * K,  # prep, cull_scatter, adjacency, vertex_update, project_escaped (rank 0 recounts below)
        "delaunay_s": wl["delaunay_s"], "sizing_s": wl["sizing_s"], "maxdp": maxdp,
    }
    if clocks is not None:
        rec["clocks"] = clocks
        rec["wall_s_timed_region"] = wall

    # ---- opt-in path (generate_mesh(ttol=...)): an iteration that re-uses the neighbour rows ----
    if full and world == 1 and loop.all_lowered:
        def reuse_step():
            loop.iterate_reuse(p_dev, p_out=p_out)

        for _ in range(3):
            reuse_step()
        r_ms = timed(reuse_step, K, flush, barrier) / K
        rec["row_reuse_step"] = {"ms_per_step": r_ms, "value": N * 1e3 / r_ms, "unit": UNIT,
                                 "what": "force iteration without retriangulation (bar pass + vertex update on the rows of the last Delaunay)"}
        loop.iterate(p_dev, t_dev, p_out=p_out)  # leave the plan as the full iteration leaves it

    # ---- sliver_removal's device work per pass (configs[1]): cull + 6 dihedral angles per kept cell + bound test
    if full and world == 1 and dim == 3:
        lo_b, hi_b = 10.0 * np.pi / 180, np.pi
        fl = torch.empty(T, dtype=torch.uint8, device=p_dev.device)
        prog0 = loop._progs[0]

        def sliver_pass():
            check(lib.dm_sliver_flags(D.ptr(prog0), D.ptr(p_dev), D.ptr(t_dev), T, geps, lo_b, hi_b, None, D.ptr(fl),
                                      D.stream_ptr()), "sliver_flags")

        for _ in range(2):
            sliver_pass()
        s_ms = timed(sliver_pass, K, flush, barrier) / K
        rec["sliver_pass"] = {"ms_per_pass": s_ms, "cells_per_s": T * 1e3 / s_ms, "slivers_flagged": int(fl.sum().item()),
                              "what": "cull + dihedral-angle bound test of one sliver_removal pass (one fused kernel)"}
        loop.iterate(p_dev, t_dev, p_out=p_out)

    if rank != 0:
        return None

    # ---- per-kernel timing (CUDA events recorded inside the library, same stream) ----
    E = loop.plan.num_bars()
    Tk = int(loop.plan.keep()[:T].sum().item())
    cap, stride = 32, 48
    acc, names = None, None
    reps = 5
    f = loop.size.struct()
    progs = D.prog_array(loop._progs)
    for r in range(reps):
        flush.fill_(r)
        ms = (C.c_float * cap)()
        nm = C.create_string_buffer(cap * stride)
        n = C.c_int(0)
        check(lib.dm_force_iteration_profiled(
            C.byref(loop.plan.c), progs, len(loop._progs), C.byref(f), D.ptr(p_dev), D.ptr(t_dev), D.ptr(p_out),
            geps, loop.L0mult, loop.delta_t, deps, h0, 0, None, D.stream_ptr(), ms, nm, stride, cap, C.byref(n)), "profiled")
        cur = np.array(ms[: n.value], dtype=np.float64)
        acc = cur if acc is None else acc + cur
        names = [nm.raw[i * stride : (i + 1) * stride].split(b"\0")[0].decode() for i in range(n.value)]
    kern_ms = acc / reps
    alg = algorithmic_bytes(n_owned if layout is not None else N, T, Tk, E, dim, grid_fh=wl["fh_grid"] is not None)
    alg = {k: v for k, v in alg.items() if k in names}
    peak, peak_src = measured_peak()
    table = []
    for nme, msv in zip(names, kern_ms):
        b = alg.get(nme, 0)
        gbs = b / (msv * 1e-3) / 1e9 if msv > 0 else 0.0
        table.append({"kernel": nme, "ms": float(msv), "share": float(msv / kern_ms.sum()), "alg_bytes": int(b),
                      "achieved_gbs": gbs, "frac": gbs / peak})
    dom_k = max(table, key=lambda r: r["ms"])
    step_bytes = sum(alg.values())
    rec["roofline"] = {
        "bound": "hbm", "kernel": dom_k["kernel"], "achieved": dom_k["achieved_gbs"], "peak": peak, "unit": "GB/s",
        "frac": dom_k["frac"], "traffic": ncu_traffic(wl["desc"], dom_k["kernel"]), "peak_source": peak_src,
        "whole_step": {"alg_bytes": int(step_bytes), "achieved": step_bytes / (kern_ms.sum() * 1e-3) / 1e9,
                       "frac": step_bytes / (kern_ms.sum() * 1e-3) / 1e9 / peak,
                       "alg_bytes_per_vertex_update": step_bytes / N},
        "kernels": [{"kernel": r["kernel"], "ms": r["ms"], "frac": r["frac"]} for r in table],
    }
    rec["config"].update({"T_kept": Tk, "bars": E})
    rec["gpu_launches"] = len(names) * K  # our kernels per step (counted from the events recorded inside the library) x steps
    if kernel_table:
        os.makedirs(os.path.dirname(os.path.abspath(kernel_table)), exist_ok=True)
        with open(kernel_table, "w") as fo:
            json.dump({"workload": wl["desc"], "h0": h0, "N": N, "T": T, "T_kept": Tk, "E": E, "peak_gbs": peak,
                       "kernels": table}, fo, indent=1)
    print(f"  [{wl['desc']}]", file=sys.stderr)
    for row in table:
        print("  %-26s %8.4f ms  %5.1f%%  %8.1f GB/s  %5.1f%% of peak" % (
            row["kernel"], row["ms"], 100 * row["share"], row["achieved_gbs"], 100 * row["frac"]), file=sys.stderr)

    # ---- CPU: the reference's loop body (oracle port) on one host core, same (p, t): baseline + parity ----
    if cpu_check and world == 1 and wl["spec"] is not None:
        step, kind = oracle_step_fn(wl["spec"], h0, wl["fh_grid"])
        nrep = 1 if N > 200000 else max(1, int(2e5 // N))
        c0 = time.perf_counter()
        for _ in range(nrep):
            ref_p = step(p, t)
        cdt = time.perf_counter() - c0
        rec["cpu_baseline"] = {"value": N * nrep / cdt, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"{nrep} step(s) of the same (p,t): N={N}, T={T}; {cdt:.1f} s", "detail": kind,
                               "max_abs_dp_vs_oracle": float(np.abs(out_pin.numpy() - ref_p).max())}
        rec["max_abs_dp_vs_oracle"] = rec["cpu_baseline"]["max_abs_dp_vs_oracle"]
    return rec


def time_to_mesh(cases, iters):
    """BASELINE.json's second metric: generate_mesh end to end -- host Delaunay every iteration included,
    termination clean-up included -- with the reference's semantics (retriangulate every iteration) and
    with the opt-in `ttol` (retriangulate when some vertex moved more than ttol local mesh sizes)."""
    import seismicmesh_b200 as sm
    from seismicmesh_b200 import meshutil

    out = {}
    for name, dom, edge, ttols in cases:
        res = {}
        for mode, kw in [("reference_semantics", {})] + [(f"ttol_{v:g}", {"ttol": v}) for v in ttols]:
            c0 = time.perf_counter()
            pm, tm = sm.generate_mesh(dom, edge, max_iter=iters, verbose=0, **kw)
            wall_m = time.perf_counter() - c0
            st_ = dict(sm.last_run_stats)
            qm = meshutil.simp_qual(pm, tm)
            res[mode] = {"wall_s": wall_m, "delaunay_s": st_["delaunay"], "device_s": st_["device"],
                         "termination_s": st_["termination"], "triangulations": st_["triangulations"],
                         "vertices": int(len(pm)), "cells": int(len(tm)), "mean_quality": float(qm.mean()),
                         "min_quality": float(qm.min()),
                         "vertex_updates_per_s_wall": st_["nverts"] * st_["iterations"] / wall_m}
        if name.startswith("ball"):
            # BASELINE.json configs[1], second half: sliver_removal on the mesh just made (the reference-
            # semantics points are regenerated: `pm` above belongs to the last ttol run)
            pm, tm = sm.generate_mesh(dom, edge, max_iter=iters, verbose=0)
            c0 = time.perf_counter()
            ps, ts = sm.sliver_removal(points=pm, domain=dom, edge_length=edge, verbose=0)
            wall_s = time.perf_counter() - c0
            ss = dict(sm.last_run_stats)
            dh = sm.geometry.calc_dihedral_angles(ps, ts)
            res["sliver_removal"] = {"wall_s": wall_s, "delaunay_s": ss.get("delaunay"), "passes": ss.get("iterations"),
                                     "vertices": int(len(ps)), "cells": int(len(ts)),
                                     "min_dihedral_deg": float(np.min(dh) * 180 / np.pi),
                                     "what": "sliver_removal(points=<the 25-iteration mesh>, min dihedral bound 10 deg, max_iter 50)"}
        res["max_iter"] = iters
        res["triangulator"] = st_["triangulator"]
        out[name] = res
    return out


def time_to_mesh_slabs(world, rank, iters):
    """Time-to-mesh on N GPUs (BASELINE.json's metric names 1/2/4/8): generate_mesh(comm=...) end to end --
    initial points, slab decomposition (decomp.blocker), per-iteration host Delaunay of every slab with the
    host cores divided among the ranks, device loop, halo migration, gather on rank 0 -- on a box of
    `world` unit cubes at h0 = 0.031 (33 k vertices per GPU, the size of the N = 1 block's ball: weak scaling)
    and on the EAGE-shaped domain at hmin 150 m / 2 Hz (a fixed mesh cut into `world` slabs: strong scaling)."""
    import torch
    import torch.distributed as dist

    import seismicmesh_b200 as sm
    from seismicmesh_b200 import meshutil
    from seismicmesh_b200.parallel import TorchComm

    comm = TorchComm()
    vp, bbox = synth_vp("eage")
    _, _, _, kw = sizing_kwargs("eage", vp, 150.0, 2.0)
    ef = sm.get_sizing_function_from_segy(None, bbox, velocity_data=vp, **kw)
    del vp
    cases = ((f"box_of_{world}_unit_cubes_h0=0.031", sm.Cube((0.0, 1.0, 0.0, float(world), 0.0, 1.0)), 0.031, 1, "weak"),
             ("eage_shaped_hmin=150_freq=2", sm.Cube(ef.bbox), ef, 1, "strong"))
    out = {}
    for name, dom, edge, axis, scaling in cases:
        dist.barrier()
        torch.cuda.synchronize()
        c0 = time.perf_counter()
        res = sm.generate_mesh(dom, edge, comm=comm, max_iter=iters, axis=axis, verbose=0)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - c0, float(sm.last_run_stats.get("delaunay", 0.0))], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            pm, tm = res
            qm = meshutil.simp_qual(pm, tm)
            out[name] = {"wall_s_max_over_ranks": float(tt[0]), "delaunay_s_max_over_ranks": float(tt[1]), "ranks": world,
                         "scaling": scaling, "vertices": int(len(pm)), "cells": int(len(tm)), "mean_quality": float(qm.mean()),
                         "min_quality": float(qm.min()), "max_iter": iters, "triangulator": sm.last_run_stats.get("triangulator")}
    return out if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ball", choices=["ball", "disk", "bp2004", "eage"])
    ap.add_argument("--h0", type=float, default=None, help="override the lattice spacing / hmin (scale-up runs)")
    ap.add_argument("--freq", type=float, default=None, help="bp2004 / eage: override the sizing frequency")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline record only: skip the bp2004 / eage sub-records and the time-to-mesh block")
    ap.add_argument("--triangulator", default=None, choices=["native", "qhull"],
                    help="host Delaunay of the set-up (untimed); it also decides the ORDER of the cell list the device "
                         "step is fed, which stage A is sensitive to.  Default: the package's (native)")
    ap.add_argument("--time-to-mesh", type=int, default=25, metavar="ITERS",
                    help="iterations of the time-to-mesh block (generate_mesh end to end, host Delaunay included)")
    ap.add_argument("--kernel-table", default=None, help="write the per-kernel roofline table (json) here")
    args = ap.parse_args()
    global TRIANGULATOR
    TRIANGULATOR = args.triangulator
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import seismicmesh_b200 as sm

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, args.warmup
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=torch.device("cuda", local_rank))  # > 126 MB L2
    default_run = args.workload == "ball" and args.h0 is None and not args.no_extras

    if world == 1:
        wl = build_workload(args.workload, args.h0, args.freq)
    else:
        wl = build_slab_workload(args.workload, rank, world, args.h0, args.freq)
    rec = measure(wl, K, W, rank, world, local_rank, True, flush, kernel_table=args.kernel_table,
                  cpu_check=not args.no_cpu_baseline)
    del wl

    extras = {}
    backend = delaunay_backend(rec["config"]["dim"]) if rank == 0 else None  # (the record is rank 0's)

    def final_line():
        return {
            "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": rec["config"], "clocks": rec.get("clocks"), "e2e": rec["e2e"],
            "gpu_launches": rec["gpu_launches"], "roofline": rec.get("roofline"), "cpu_baseline": rec.get("cpu_baseline"),
            "row_reuse_step": rec.get("row_reuse_step"), "sliver_pass": rec.get("sliver_pass"),
            "workloads": extras.get("workloads"), "time_to_mesh": extras.get("time_to_mesh"),
            "delaunay_s": rec["delaunay_s"], "delaunay_backend": backend, "sizing_s": rec["sizing_s"],
            "maxdp": rec["maxdp"], "wall_s_timed_region": rec.get("wall_s_timed_region"),
        }

    # The sub-records of a multi-rank run go through collectives: should one rank fail inside them the others
    # would wait for ever and the headline line (measured above) would be lost with them.  A watchdog prints
    # the line without the missing sub-records and ends the process instead.
    watchdog = None
    if default_run and world > 1:
        def bail():
            if rank == 0:
                extras.setdefault("workloads", {"error": "sub-records timed out"})
                extras.setdefault("time_to_mesh", {"error": "timed out"})
                print(json.dumps(final_line()), flush=True)
            os._exit(0)

        watchdog = threading.Timer(float(os.environ.get("DM_BENCH_SUBRECORD_TIMEOUT", "420")), bail)
        watchdog.daemon = True
        watchdog.start()
    if default_run and world == 1:
        subs = {}
        for name, wk, h, fq in (("bp2004_hmin25_freq6", "bp2004", 25.0, 6.0), ("eage_hmin75_freq4", "eage", 75.0, 4.0)):
            w2 = build_workload(wk, h, fq)
            r2 = measure(w2, max(5, K // 2), 3, rank, world, local_rank, False, flush)
            del w2
            torch.cuda.empty_cache()
            subs[name] = {k: r2[k] for k in ("value", "unit", "ms_per_step", "config", "e2e", "roofline", "max_abs_dp_vs_oracle",
                                             "delaunay_s", "sizing_s") if k in r2}
        extras["workloads"] = subs
        if args.time_to_mesh > 1:
            # ttol values: 0.1 is DistMesh's 2-D default; in 3-D the maximum over all vertices is carried by a
            # few boundary vertices and 0.5 local mesh sizes is what skips retriangulations at equal quality
            # (profiles/r2g_ttol_sweep.json)
            vp, bbox = synth_vp("eage")
            _, _, _, kw = sizing_kwargs("eage", vp, 150.0, 2.0)
            ef = sm.get_sizing_function_from_segy(None, bbox, velocity_data=vp, **kw)
            del vp
            extras["time_to_mesh"] = time_to_mesh(
                (("disk_h0=0.01", sm.Disk([0.0, 0.0], 1.0), 0.01, (0.1, 0.3)),
                 ("ball_h0=0.05", sm.Ball([0.0, 0.0, 0.0], 1.0), 0.05, (0.1, 0.5)),
                 # (the per-GPU workload of the N > 1 block's weak-scaling case: box of N unit cubes)
                 ("box_of_1_unit_cubes_h0=0.031", sm.Cube((0.0, 1.0, 0.0, 1.0, 0.0, 1.0)), 0.031, ()),
                 ("eage_shaped_hmin=150_freq=2", sm.Cube(ef.bbox), ef, (0.5,))),
                args.time_to_mesh)
            del ef
            torch.cuda.empty_cache()
    elif default_run and world > 1:
        try:
            w2 = build_slab_workload("eage", rank, world)
            r2 = measure(w2, max(5, K // 2), 3, rank, world, local_rank, False, flush)
            if rank == 0:
                extras["workloads"] = {"eage_slabs": {k: r2[k] for k in ("value", "unit", "ms_per_step", "config", "e2e", "roofline",
                                                                          "delaunay_s", "sizing_s") if k in r2}}
        except Exception as exc:  # never lose the headline line to a sub-record
            if rank == 0:
                extras["workloads"] = {"eage_slabs": {"error": f"{type(exc).__name__}: {exc}"}}
        if args.time_to_mesh > 1:
            try:
                extras["time_to_mesh"] = time_to_mesh_slabs(world, rank, args.time_to_mesh)
            except Exception as exc:
                if rank == 0:
                    extras["time_to_mesh"] = {"error": f"{type(exc).__name__}: {exc}"}

    if watchdog is not None:
        watchdog.cancel()
    if rank == 0:
        print(json.dumps(final_line()), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
