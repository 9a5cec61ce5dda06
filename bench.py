#!/usr/bin/env python
"""bench.py -- DistMesh force-iteration throughput (vertex-updates/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ball|disk|bp2004|eage] [--impl ours|reference]

A *step* is one pass of the hot path (cull -> unique bars -> forces -> update -> projection ->
max|dp|, i.e. one iteration of the loop body of SeismicMesh generate_mesh AFTER its Delaunay
call) over one synthetic (p, t).  Delaunay is host work in the reference's design and is done
once in the untimed set-up; it is reported separately (`delaunay_s`).

Prints ONE JSON line (contract in the task statement): `value` = vertex-updates/s with inputs
resident in HBM, `e2e` = the same through the public host-buffer call (H2D of p and t from pinned
memory + D2H of the new positions inside the timed region), `roofline` for the dominant kernel
(algorithmic bytes / CUDA-event time against MEASURED_PEAKS.json), `cpu_baseline` = the oracle
port of the reference's NumPy loop body timed on one host core.
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


# ------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# ------------------------------------------------------------------------------------------------
def _lattice(h0, dim, bbox):
    axes = [np.arange(int(np.ceil((hi + h0 - lo) / h0)), dtype=float) * h0 + lo for lo, hi in bbox]
    g = [a.copy() for a in np.meshgrid(*axes, indexing="ij")]
    g[1][1::2] += h0 / 2
    if dim == 3:
        g[2][1::2] += h0 / 2
    return np.stack([a.ravel() for a in g], axis=1)


def make_points(workload, h0, seed=0, shift=0.0):
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
    p[:, 1] += shift
    return np.ascontiguousarray(p), dim


TRIANGULATOR = None  # --triangulator: None = the package default (native), "qhull", "native"


def delaunay_backend(dim):
    from seismicmesh_b200.triangulator import get_triangulator

    return get_triangulator(TRIANGULATOR, dim).name


def triangulate(p):
    """Host Delaunay with the package's default triangulator (2-D: native sweep-hull, 3-D: Qhull).
    DM_BENCH_CACHE=<dir> re-uses the cells of an identical point set between runs of one profiling
    session (the set-up is untimed either way)."""
    from seismicmesh_b200.triangulator import get_triangulator

    cache = os.environ.get("DM_BENCH_CACHE")
    key = None
    if cache:
        import hashlib

        key = os.path.join(cache, f"tri_{TRIANGULATOR or 'default'}_" + hashlib.sha1(p.tobytes()).hexdigest()[:16] + ".npz")
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
        spec = ("ball", dict(x0=[0.0, 0.0, 0.0], r=1.0)) if dim == 3 else ("disk", dict(x0=[0.0, 0.0], r=1.0))
        return dict(p=p, t=t, dim=dim, dom=dom, size=SizeSpec(dim, const=h0), h0=h0, spec=spec, fh_grid=None,
                    delaunay_s=dt, desc=f"{workload}_h0={h0:g}", sizing_s=0.0, edge=h0)
    import torch

    from seismicmesh_b200 import device as D
    from seismicmesh_b200.engine import ForceLoop, Level
    from seismicmesh_b200.generation import _initial_points

    vp, bbox = synth_vp(workload)
    if workload == "bp2004":  # README.md:176-186, benchmarks/benchmark_BP2004.py:31-39
        hmin, fr = h0 or 75.0, freq or 2.0
        kw = dict(hmin=hmin, wl=10, freq=fr, dt=0.001, grade=0.15, domain_pad=1e3, pad_style="edge",
                  nz=vp.shape[0], nx=vp.shape[1])
        dim = 2
    else:  # README.md:275-292
        hmin, fr = h0 or 150.0, freq or 2.0
        kw = dict(hmin=hmin, wl=5, freq=fr, dt=0.001, grade=0.15, hmax=5e3, domain_pad=250.0,
                  pad_style="linear_ramp", nz=vp.shape[0], nx=vp.shape[1], ny=vp.shape[2])
        dim = 3
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
    spec = ("rectangle", dict(bbox=tuple(ef.bbox))) if dim == 2 else ("cube", dict(bbox=tuple(ef.bbox)))
    g = ef.interpolant()
    return dict(p=p, t=t, dim=dim, dom=dom, size=size, h0=hmin, spec=spec, fh_grid=(g.grid, g.values),
                delaunay_s=dt, sizing_s=sizing_s, edge=ef,
                desc=f"{workload}_shaped_grid{'x'.join(str(n) for n in g.values.shape)}_hmin={hmin:g}_freq={fr:g}")


def oracle_step_fn(wl):
    """The reference's loop body (oracle port, + the reference's own native unique_edges from
    oracle/_ref when it was built) as a closure step(p, t) -> p_new."""
    from oracle import distmesh_oracle as orc
    from oracle import ref_harness

    spec, h0 = wl["spec"], wl["h0"]
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
    if wl["fh_grid"] is None:
        fh = lambda x: np.array([h0] * len(x))  # noqa: E731
    else:
        axes, grid = wl["fh_grid"]
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
    hs = 8 * E if grid_fh else 0
    return {
        # zero the per-iteration counters; 3-D: p read once, the padded copy written once
        "prep(zero+pad)": 4 * (N + 1) + (56 * N if dim == 3 else 0),
        # t and p read once, keep flags + the kept cells handed to the bar stage written once
        "cull_scatter": 4 * c * T + 8 * d * N + T + 4 * c * Tk,
        # SURVEY 8d K2 (+K5: rows are symmetric, every bar is stored at both of its ends) with the
        # bar pass K3 fused in (positions read once, h written once per bar for gridded fh)
        "adjacency": 4 * c * Tk + 8 * (N + 1) + 8 * E + 8 * d * N + hs,
        "bar_pass+scale": 8 * d * N + 4 * E + 8 * N + hs,
        "vertex_update+maxdp": 16 * d * N + 8 * E + 8 * N + 2 * hs,
    }


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json); None if there is none."""
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


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path, on the host cores of this box."""
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    workload = args.workload
    if workload in ("ball", "disk"):
        base_h0 = args.h0 or (0.02 if workload == "ball" else 0.01)
        # bounded sample: coarsen h0 until (K+W) steps + set-up fit in ~150 s (cost ~ 1/h0^dim)
        dim = 3 if workload == "ball" else 2
        est_full = 8.0 if workload == "ball" else 0.12
        h0 = base_h0
        for cand in (1.0, 1.25, 1.5, 2.0, 3.0):
            h0 = base_h0 * cand
            if (K + W) * est_full / cand**dim + (30.0 if workload == "ball" else 1.0) / cand**dim <= 150.0:
                break
        wl = build_workload(workload, h0)
    else:
        # gridded workloads need the device for their set-up (sizing function, settling iterations);
        # the timed loop body below is host-only
        wl = build_workload(workload, args.h0, args.freq)
        base_h0 = h0 = wl["h0"]
    p, t, dim, tq = wl["p"], wl["t"], wl["dim"], wl["delaunay_s"]
    step, kind = oracle_step_fn(wl)
    for _ in range(W):
        step(p, t)
    t0 = time.perf_counter()
    for _ in range(K):
        step(p, t)
    dt = time.perf_counter() - t0
    N = len(p)
    val = N * K / dt
    sample = f"{wl['desc']} (N={N}, T={len(t)}), {K} steps of the full loop body on 1 core"
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": wl["desc"] if workload not in ("ball", "disk") else f"{workload}_h0={base_h0:g}", "sample_h0": h0, "delaunay": "excluded (host, set-up)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "detail": kind},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "delaunay_s": tq,
    }
    print(json.dumps(out), flush=True)


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
    ap.add_argument("--triangulator", default=None, choices=["native", "qhull"],
                    help="host Delaunay of the set-up (untimed); it also decides the ORDER of the cell list the device "
                         "step is fed, which stage A is sensitive to.  Default: the package's (native)")
    ap.add_argument("--time-to-mesh", type=int, default=0, metavar="ITERS",
                    help="also run generate_mesh(max_iter=ITERS) end to end (host Delaunay included) and report "
                         "time-to-mesh, with the reference's retriangulate-every-iteration and with ttol=0.1")
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
    from seismicmesh_b200 import device as D
    from seismicmesh_b200._lib import check, lib
    from seismicmesh_b200.engine import ForceLoop, Level, SizeSpec

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    workload = args.workload
    K, W = args.steps, args.warmup

    # ---- set-up (untimed): points + ONE host Delaunay per rank (weak scaling: one slab per GPU) ----
    layout = None
    if world == 1:
        wl = build_workload(workload, args.h0, args.freq)
        p, t, dim, dom, h0, t_delaunay = wl["p"], wl["t"], wl["dim"], wl["dom"], wl["h0"], wl["delaunay_s"]
        size = wl["size"]
        n_owned = len(p)
    else:
        # weak scaling: a cylinder (3-D) / rectangle (2-D) of `world` slabs along axis 1, each slab
        # with the volume of the N=1 ball / disk; every rank meshes its slab + ghost layers
        from seismicmesh_b200.parallel import make_slab_workload

        if workload not in ("ball", "disk"):
            raise SystemExit("multi-GPU bench: --workload ball|disk (slab workload)")
        h0 = args.h0 or (0.02 if workload == "ball" else 0.01)
        p, dim, dom, layout = make_slab_workload(workload, h0, rank, world)
        n_owned = layout.n_owned
        t, t_delaunay = triangulate(p)
        t = np.ascontiguousarray(t[(t < n_owned).any(axis=1)])  # cells made only of ghosts belong to the neighbours
        size = SizeSpec(dim, const=h0)
        wl = dict(desc=f"{workload}_h0={h0:g}", fh_grid=None, sizing_s=0.0)
    N, T = len(p), len(t)
    geps, deps = 0.1 * h0, np.sqrt(np.finfo(np.double).eps) * h0
    loop = ForceLoop(dim, [Level(dom, dim)], size, h0, geps, deps)

    p_dev = D.to_dev(p, torch.float64)
    t_dev = D.to_dev(t, torch.int32)
    p_out = torch.empty_like(p_dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=p_dev.device)  # > 126 MB L2

    halo, halo_kind = None, None
    if world > 1:
        from seismicmesh_b200.parallel import PeerHalo, RingHalo

        ring = RingHalo(layout, dim, p_dev.device, rank=rank, world=world)
        halo, halo_kind = ring, "NCCL P2P send/recv (RingHalo)"
        if os.environ.get("DM_HALO", "p2p") == "p2p":
            try:  # NVLink peer-memory push with our own kernel; checked once against the NCCL exchange
                peer = PeerHalo(layout, dim, p_dev.device, rank=rank, world=world)
                a, b = p_dev.clone(), p_dev.clone()
                a[n_owned:] = float("nan")
                b[n_owned:] = float("nan")
                ring.exchange(a)
                peer.exchange(b)
                peer.exchange(b)  # both slots
                torch.cuda.synchronize()
                assert torch.equal(a, b), "PeerHalo and RingHalo disagree"
                halo, halo_kind = peer, "NVLink peer-memory push, dm_halo_push + symmetric-memory signals (PeerHalo)"
            except Exception as exc:  # no peer access / symmetric memory on this box: keep NCCL
                halo_kind += f" [peer-memory path unavailable: {type(exc).__name__}: {exc}]"

    def one_step():
        loop.iterate(p_dev, t_dev, p_out=p_out)
        if halo is not None:
            halo.exchange(p_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # pinned host buffers of the e2e leg (allocated before the clocks sampler starts)
    p_pin = torch.from_numpy(p).pin_memory()
    t_pin = torch.from_numpy(t).pin_memory()
    out_pin = torch.empty_like(p_pin).pin_memory()
    sc_pin = torch.empty(8, dtype=torch.float64).pin_memory()
    for _ in range(W):
        one_step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    wall0 = time.perf_counter()
    for i in range(K):
        flush.fill_(i & 0xFF)  # L2 flush between timed iterations (not timed)
        ev[i][0].record()
        one_step()
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(step_ms.sum())
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=p_dev.device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
        nn = torch.tensor([n_owned], dtype=torch.float64, device=p_dev.device)
        dist.all_reduce(nn, op=dist.ReduceOp.SUM)
        N_all = int(nn.item())
    else:
        N_all = n_owned
    value = N_all * K / (total_ms * 1e-3)

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region ----

    def e2e_step():
        # the public host-buffer call: H2D of p and t (t in chunks, stage A overlapped with the transfer),
        # stages B-D, D2H of the new positions -- all inside the timed region
        loop.iterate_host(p_pin, t_pin, out_pin)
        if halo is not None:
            halo.exchange(loop._host_bufs[2])
        sc_pin.copy_(loop.plan.scalars(), non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        flush.fill_(i & 0xFF)
        ev2[i][0].record()
        e2e_step()
        ev2[i][1].record()
    barrier()
    e2e_ms = float(sum(a.elapsed_time(b) for a, b in ev2))
    if world > 1:
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=p_dev.device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = N_all * K / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None  # sampled every 20 ms across both timed regions
    maxdp = float(sc_pin[4])

    # ---- opt-in path (generate_mesh(ttol=...)): an iteration that re-uses the neighbour rows ----
    reuse = None
    if world == 1 and loop.all_lowered:
        for _ in range(3):
            loop.iterate_reuse(p_dev, p_out=p_out)
        barrier()
        ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            flush.fill_(i & 0xFF)
            ev3[i][0].record()
            loop.iterate_reuse(p_dev, p_out=p_out)
            ev3[i][1].record()
        barrier()
        r_ms = float(sum(a.elapsed_time(b) for a, b in ev3)) / K
        reuse = {"ms_per_step": r_ms, "value": N * 1e3 / r_ms, "unit": UNIT,
                 "what": "force iteration without retriangulation (bar pass + vertex update on the rows of the last Delaunay)"}
        loop.iterate(p_dev, t_dev, p_out=p_out)  # leave the plan as the full iteration leaves it

    # ---- sliver_removal's device work per pass (configs[1]): cull + order-preserving compaction +
    #      6 dihedral angles per kept cell + bound test (mesh_generator.py:204-243)
    sliver = None
    if world == 1 and dim == 3:
        lo_b, hi_b = 10.0 * np.pi / 180, np.pi

        fl = torch.empty(T, dtype=torch.uint8, device=p_dev.device)
        prog0 = loop._progs[0]

        def sliver_pass():
            check(lib.dm_sliver_flags(D.ptr(prog0), D.ptr(p_dev), D.ptr(t_dev), T, geps, lo_b, hi_b, None, D.ptr(fl),
                                      D.stream_ptr()), "sliver_flags")
            return fl

        for _ in range(2):
            sliver_pass()
        barrier()
        ev4 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            flush.fill_(i & 0xFF)
            ev4[i][0].record()
            fl = sliver_pass()
            ev4[i][1].record()
        barrier()
        s_ms = float(sum(a.elapsed_time(b) for a, b in ev4)) / K
        sliver = {"ms_per_pass": s_ms, "cells_per_s": T * 1e3 / s_ms, "slivers_flagged": int(fl.sum().item()),
                  "what": "cull + dihedral-angle bound test of one sliver_removal pass (one fused kernel)"}
        loop.iterate(p_dev, t_dev, p_out=p_out)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

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
    alg = algorithmic_bytes(N, T, Tk, E, dim, grid_fh=wl["fh_grid"] is not None)
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
    roofline = {
        "bound": "hbm", "kernel": dom_k["kernel"], "achieved": dom_k["achieved_gbs"], "peak": peak, "unit": "GB/s",
        "frac": dom_k["frac"], "traffic": ncu_traffic(wl["desc"], dom_k["kernel"]), "peak_source": peak_src,
        "whole_step": {"alg_bytes": int(step_bytes), "achieved": step_bytes / (kern_ms.sum() * 1e-3) / 1e9,
                       "frac": step_bytes / (kern_ms.sum() * 1e-3) / 1e9 / peak,
                       "alg_bytes_per_vertex_update": step_bytes / N},
    }
    if args.kernel_table:
        os.makedirs(os.path.dirname(os.path.abspath(args.kernel_table)), exist_ok=True)
        with open(args.kernel_table, "w") as fo:
            json.dump({"workload": wl["desc"], "h0": h0, "N": N, "T": T, "T_kept": Tk, "E": E, "peak_gbs": peak,
                       "kernels": table}, fo, indent=1)
    for row in table:
        print("  %-26s %8.4f ms  %5.1f%%  %8.1f GB/s  %5.1f%% of peak" % (
            row["kernel"], row["ms"], 100 * row["share"], row["achieved_gbs"], 100 * row["frac"]), file=sys.stderr)

    # ---- CPU baseline: the reference's loop body (oracle port) on one host core, same (p, t) ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        step, kind = oracle_step_fn(wl)
        p0, t0_ = p, t
        nrep = 1 if N > 200000 else max(1, int(2e5 // N))
        c0 = time.perf_counter()
        for _ in range(nrep):
            ref_p = step(p0, t0_)
        cdt = time.perf_counter() - c0
        cpu = {"value": len(p0) * nrep / cdt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{nrep} step(s) of the same (p,t): N={len(p0)}, T={len(t0_)}; {cdt:.1f} s", "detail": kind}
        if world == 1:  # parity of the benchmarked step against the oracle
            perr = float(np.abs(out_pin.numpy() - ref_p).max())
            cpu["max_abs_dp_vs_oracle"] = perr

    ttm = None
    if args.time_to_mesh > 0 and world == 1:
        from seismicmesh_b200 import meshutil

        edge = wl["edge"]
        ttm = {}
        for name, kw in (("reference_semantics", {}), ("ttol_0.1", {"ttol": 0.1})):
            c0 = time.perf_counter()
            pm, tm = sm.generate_mesh(dom, edge, max_iter=args.time_to_mesh, verbose=0, **kw)
            wall_m = time.perf_counter() - c0
            st_ = dict(sm.last_run_stats)
            qm = meshutil.simp_qual(pm, tm)
            ttm[name] = {"wall_s": wall_m, "delaunay_s": st_["delaunay"], "device_s": st_["device"],
                         "triangulations": st_["triangulations"], "vertices": int(len(pm)), "cells": int(len(tm)),
                         "mean_quality": float(qm.mean()), "min_quality": float(qm.min())}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl["desc"], "N_per_gpu": N, "T_per_gpu": T, "T_kept": Tk, "bars": E, "dim": dim,
                   "l2": "flushed between timed steps (512 MiB fill, untimed)", "delaunay": "host, set-up (untimed)",
                   "parallelism": (f"{world} slabs along axis 1 (owned+ghost per GPU), halo exchange per step: {halo_kind}; "
                                   f"halo bytes/step/rank={halo.bytes_per_exchange}") if world > 1 else "single",
                   "N_owned_total": N_all},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(p.nbytes + t.nbytes),
                "d2h_bytes_per_step": int(p.nbytes + 64), "ms_per_step": e2e_ms / K},
        "gpu_launches": 5 * K,  # prep, cull_scatter, adjacency, vertex_update, project_escaped
        "roofline": roofline, "cpu_baseline": cpu, "row_reuse_step": reuse, "sliver_pass": sliver, "time_to_mesh": ttm,
        "delaunay_s": t_delaunay, "delaunay_backend": delaunay_backend(dim), "sizing_s": wl["sizing_s"], "maxdp": maxdp,
        "wall_s_timed_region": wall,
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
