"""Device-resident DistMesh force iteration: the thin host driver over the C ABI.

One :class:`ForceLoop` owns the per-iteration workspace (a torch uint8 blob carved by
``dm_plan_init``) and runs stages A-D of include/distmesh_b200.h on the current CUDA stream:

    A  cull (fd on centroids) + raw-bar counting      mesh_generator.py:734-738
    B  unique bars (bit-exact) + lower-neighbour CSR   mesh_generator.py:680-688, fast_geometry.cpp:30-77
    C  L, fh(midpoints), global force scale            mesh_generator.py:696-700
    D  force gather, pfix, update, projection, maxdp   mesh_generator.py:701-712, 499-521

``fd`` levels and ``fh`` may each be *lowered* (geometry objects / gridded or scalar sizing) or
*opaque* Python callables; opaque ones are user code and are evaluated on the host through
staged copies (their time is reported as host time), everything else stays on the device.
"""
import ctypes as C
import os
import time

import numpy as np
import torch

from . import _lib
from . import device as D
from ._lib import check, lib
from .geometry import lower


class SizeSpec:
    """fh in one of three forms: constant, device grid, opaque callable."""

    def __init__(self, dim, const=None, interp=None, func=None):
        self.dim = dim
        self.const = const
        self.interp = interp
        self.func = func

    @property
    def lowered(self):
        return self.func is None

    def struct(self):
        if self.const is not None:
            return D.size_fn_struct(_lib.SIZE_CONST, self.dim, hconst=self.const)
        if self.interp is not None:
            return self.interp.struct()
        return D.size_fn_struct(_lib.SIZE_EXTERNAL, self.dim)

    def eval_host(self, x):
        """fh at host points x -> host array (used for the initial points only)."""
        if self.const is not None:
            return np.array([self.const] * len(x))
        if self.interp is not None:
            return self.interp(x)
        return np.asarray(self.func(x), dtype=np.float64)


class Level:
    """fd level set: lowered device program or opaque callable."""

    def __init__(self, domain_or_callable, dim):
        self.dim = dim
        self.obj = domain_or_callable
        self.prog = lower(domain_or_callable)
        if self.prog is None:
            self.func = domain_or_callable.eval if hasattr(domain_or_callable, "eval") else domain_or_callable
        else:
            self.func = None

    @property
    def lowered(self):
        return self.prog is not None

    def eval_host(self, x):
        if self.prog is not None:
            return self.obj.eval(x)
        return np.asarray(self.func(x), dtype=np.float64)


def _project_host(p, fd, deps, hmin, idx):
    """Projection for an OPAQUE fd (user code on the host), semantics of
    _project_points_back_newton (mesh_generator.py:762-784)."""
    dim = p.shape[1]
    d = fd(p)
    ix = d > 0.0 if idx == 0 else np.logical_and(d > 0.0, d < hmin / 1.5)
    if ix.any():
        grads = []
        for i in range(dim):
            dv = np.zeros(dim)
            dv[i] = deps
            grads.append((fd(p[ix] + dv) - d[ix]) / deps)
        g2 = sum(g**2 for g in grads)
        g2 = np.where(g2 < deps, deps, g2)
        p[ix] -= (d[ix] * np.vstack(grads) / g2).T
    return p


LAYOUTS = {"auto": 0, "buckets": 1, "tiles": 2}  # DM_LAYOUT_* of include/distmesh_b200.h


class ForceLoop:
    def __init__(self, dim, levels, size, h0, geps, deps, delta_t=0.30, nfix=0, layout=None):
        """layout: how stages A + B pass the kept cells to the vertices ("buckets" / "tiles" / "auto"; same
        rows either way).  None: DM_LAYOUT from the environment, else the measured choice -- per-vertex
        buckets with a 3-D gridded fh (its bar pass wants many short-lived warps), otherwise auto (the
        library takes the tile layout from 200 000 rows on)."""
        D.require_cuda()
        if layout is None:
            layout = os.environ.get("DM_LAYOUT") or ("buckets" if dim == 3 and size.interp is not None else "auto")
        if layout not in LAYOUTS:
            raise ValueError(f"layout must be one of {sorted(LAYOUTS)}")
        self.layout = layout
        self.dim = dim
        self.levels = levels
        self.size = size
        self.h0, self.geps, self.deps, self.delta_t = float(h0), float(geps), float(deps), float(delta_t)
        self.L0mult = 1 + 0.4 / 2 ** (dim - 1)
        self.nfix = int(nfix)
        self.fixed_mask = None  # optional (N,) uint8 device mask of additional fixed vertices
        self.n_rows = None      # slabs: local vertices are [owned | ghosts] and only the first n_rows are owned
        self.plan = None
        self.all_lowered = all(lv.lowered for lv in levels) and size.lowered
        self._progs = [lv.prog for lv in levels if lv.lowered]
        self.host_seconds = 0.0  # time spent in opaque user callables + their staging copies
        self.kernel_launches = 0

    # -- workspace ------------------------------------------------------------------------
    def ensure_plan(self, N, T):
        if self.plan is None or self.plan.N != N or self.plan.T < T:
            cap = T if self.plan is None or self.plan.N != N else max(T, int(self.plan.T * 1.25))
            self.plan = D.Plan(N, cap, self.dim)
        # the C plan is sized for capacity; the live cell count is set per call
        self.plan.c.T = T
        self.plan.c.K = self.dim * (self.dim + 1) * T
        # multi-GPU slabs: only the first n_rows (owned) vertices get rows / forces / an update
        check(lib.dm_plan_set_rows(C.byref(self.plan.c), N if self.n_rows is None else int(self.n_rows)), "dm_plan_set_rows")
        check(lib.dm_plan_set_layout(C.byref(self.plan.c), LAYOUTS[self.layout]), "dm_plan_set_layout")
        return self.plan

    # -- the iteration --------------------------------------------------------------------
    def cull(self, p, t):
        """Stage A.  Returns nothing; plan.keep holds the flags."""
        pl = self.plan
        st = D.stream_ptr()
        lv0 = self.levels[0]
        if lv0.lowered:
            check(lib.dm_stage_cull_count(C.byref(pl.c), D.ptr(lv0.prog), D.ptr(p), D.ptr(t), self.geps, 1, st), "cull_count")
        else:
            t0 = time.perf_counter()
            cen = torch.empty((t.shape[0], self.dim), dtype=torch.float64, device=p.device)
            check(lib.dm_centroids(D.ptr(p), D.ptr(t), t.shape[0], self.dim, D.ptr(cen), st), "centroids")
            keep = np.asarray(lv0.func(cen.cpu().numpy())) < -self.geps
            pl.keep()[: t.shape[0]].copy_(torch.from_numpy(keep.astype(np.uint8)))
            self.host_seconds += time.perf_counter() - t0
            check(lib.dm_stage_cull_count(C.byref(pl.c), None, D.ptr(p), D.ptr(t), self.geps, 1, st), "cull_count")

    def iterate(self, p, t, p_out=None, want_forces=False):
        """One force iteration on device tensors p (N,dim) f64, t (T,dim+1) i32.
        Returns (p_new, Ftot|None).  maxdp is left in plan.scalars()[4]."""
        N, T = p.shape[0], t.shape[0]
        pl = self.ensure_plan(N, T)
        st = D.stream_ptr()
        if p_out is None:
            p_out = torch.empty_like(p)
        Ftot = torch.empty_like(p) if want_forces else None
        if self.all_lowered:
            f = self.size.struct()
            progs = D.prog_array(self._progs)
            check(
                lib.dm_force_iteration(
                    C.byref(pl.c), progs, len(self._progs), C.byref(f), D.ptr(p), D.ptr(t), D.ptr(p_out),
                    self.geps, self.L0mult, self.delta_t, self.deps, self.h0, self.nfix, D.ptr(self.fixed_mask),
                    D.ptr(Ftot), st,
                ),
                "dm_force_iteration",
            )
            return p_out, Ftot
        # ---- staged path: at least one opaque user callable ----
        self.cull(p, t)
        check(lib.dm_stage_build_adjacency(C.byref(pl.c), st), "build_bars")
        f = self.size.struct()
        if not self.size.lowered:
            t0 = time.perf_counter()
            check(lib.dm_stage_bar_index(C.byref(pl.c), st), "bar_index")
            E = pl.num_bars()
            mid = torch.empty((E, self.dim), dtype=torch.float64, device=p.device)
            check(lib.dm_bar_midpoints(C.byref(pl.c), D.ptr(p), D.ptr(mid), st), "bar_midpoints")
            h = np.ascontiguousarray(self.size.func(mid.cpu().numpy()), dtype=np.float64)
            pl.hbar(E).copy_(torch.from_numpy(h))
            self.host_seconds += time.perf_counter() - t0
        check(lib.dm_stage_bar_pass(C.byref(pl.c), D.ptr(p), C.byref(f), st), "bar_pass")
        fused = all(lv.lowered for lv in self.levels)
        progs = D.prog_array(self._progs if fused else [])
        check(
            lib.dm_stage_vertex_update(
                C.byref(pl.c), D.ptr(p), D.ptr(p_out), progs, len(self._progs) if fused else 0, C.byref(f), self.L0mult,
                self.delta_t, self.deps, self.h0, self.nfix, D.ptr(self.fixed_mask), D.ptr(Ftot), st,
            ),
            "vertex_update",
        )
        if not fused:
            for idx, lv in enumerate(self.levels):
                if lv.lowered:
                    check(lib.dm_project_points(D.ptr(lv.prog), D.ptr(p_out), N, self.dim, self.deps, self.h0, idx, st), "project")
                else:
                    t0 = time.perf_counter()
                    ph = _project_host(p_out.cpu().numpy(), lv.func, self.deps, self.h0, idx)
                    p_out.copy_(torch.from_numpy(ph))
                    self.host_seconds += time.perf_counter() - t0
        return p_out, Ftot

    def iterate_host(self, p_pin, t_pin, out_pin, chunks=8, p_dev=None):
        """One force iteration with HOST buffers in and out (pinned torch CPU tensors: p (N,dim) f64,
        t (T,dim+1) i32, out (N,dim) f64): what a driver whose Delaunay lives on the host does every
        iteration, and what `generate_mesh` itself calls.  The cell list is uploaded in `chunks` pieces
        on a copy stream and stage A runs on each piece as it lands, so the cull + scatter hides behind
        the PCIe transfer; stages B-D follow, then the new positions go back (asynchronously: the
        caller synchronises before reading `out_pin`).  `p_dev`: the positions are already on the
        device (the previous call's result) -- `p_pin` is then ignored and only the cells travel.
        Returns the device tensor of the new positions (valid until the call after the next)."""
        if not self.all_lowered:
            raise RuntimeError("iterate_host needs lowered fd / fh")
        N, T = out_pin.shape[0], t_pin.shape[0]
        pl = self.ensure_plan(N, T)
        dev = D.device()
        hb = getattr(self, "_host_bufs", None)
        if hb is None or hb[0].shape[0] != N or hb[1].shape[0] < T:
            cap = max(T, 1) if hb is None or hb[0].shape[0] != N else max(T, int(hb[1].shape[0] * 1.25))
            hb = [torch.empty((N, self.dim), dtype=torch.float64, device=dev),
                  torch.empty((cap, self.dim + 1), dtype=torch.int32, device=dev),
                  torch.empty((N, self.dim), dtype=torch.float64, device=dev), torch.cuda.Stream(device=dev),
                  torch.empty((N, self.dim), dtype=torch.float64, device=dev)]
            self._host_bufs = hb
        t_dev, cs = hb[1], hb[3]
        # two result buffers, alternating: the caller may hand the previous result back as `p_dev`
        p_out = hb[2] if (p_dev is None or p_dev.data_ptr() != hb[2].data_ptr()) else hb[4]
        p_in = hb[0] if p_dev is None else p_dev
        ms = torch.cuda.current_stream()
        st = D.stream_ptr()
        cs.wait_stream(ms)  # the previous call's kernels are done with p_in / t_dev
        nch = max(1, min(int(chunks), T)) if T > 0 else 0
        bounds = [T * k // nch for k in range(nch + 1)] if nch else [0]
        evs = []
        with torch.cuda.stream(cs):
            ev_p = None
            if p_dev is None:
                p_in.copy_(p_pin, non_blocking=True)
                ev_p = torch.cuda.Event()
                ev_p.record(cs)
            for k in range(nch):
                a, b = bounds[k], bounds[k + 1]
                t_dev[a:b].copy_(t_pin[a:b], non_blocking=True)
                e = torch.cuda.Event()
                e.record(cs)
                evs.append(e)
        if ev_p is not None:
            ms.wait_event(ev_p)
        check(lib.dm_stage_prep(C.byref(pl.c), D.ptr(p_in), st), "dm_stage_prep")
        prog0 = D.ptr(self._progs[0])
        for k in range(nch):
            a, b = bounds[k], bounds[k + 1]
            ms.wait_event(evs[k])
            check(lib.dm_stage_cull_chunk(C.byref(pl.c), prog0, D.ptr(p_in), C.c_void_p(t_dev.data_ptr() + 4 * (self.dim + 1) * a),
                                          a, b - a, self.geps, 1, st), "dm_stage_cull_chunk")
        f = self.size.struct()
        progs = D.prog_array(self._progs)
        check(
            lib.dm_force_iteration_tail(
                C.byref(pl.c), progs, len(self._progs), C.byref(f), D.ptr(p_in), D.ptr(p_out), self.L0mult, self.delta_t,
                self.deps, self.h0, self.nfix, D.ptr(self.fixed_mask), None, st,
            ),
            "dm_force_iteration_tail",
        )
        out_pin.copy_(p_out, non_blocking=True)
        return p_out

    def iterate_reuse(self, p, p_out=None, want_forces=False):
        """One force iteration WITHOUT retriangulation: stages C + D on the neighbour rows left by the
        last :meth:`iterate` (same cell list, moved points).  Opt-in (`ttol`); lowered fd / fh only."""
        if not self.all_lowered or self.plan is None or self.plan.N != p.shape[0]:
            raise RuntimeError("iterate_reuse needs a previous iterate() on the same vertices and lowered fd / fh")
        pl = self.plan
        if p_out is None:
            p_out = torch.empty_like(p)
        Ftot = torch.empty_like(p) if want_forces else None
        f = self.size.struct()
        progs = D.prog_array(self._progs)
        check(
            lib.dm_force_iteration_reuse(
                C.byref(pl.c), progs, len(self._progs), C.byref(f), D.ptr(p), D.ptr(p_out), self.L0mult, self.delta_t,
                self.deps, self.h0, self.nfix, D.ptr(self.fixed_mask), D.ptr(Ftot), D.stream_ptr(),
            ),
            "dm_force_iteration_reuse",
        )
        return p_out, Ftot

    def spare_like(self, p):
        """An (N,dim) device buffer that is not `p` (two alternate; for ping-pong callers)."""
        sp = getattr(self, "_spare", None)
        if sp is None or sp[0].shape != p.shape:
            sp = [torch.empty_like(p), torch.empty_like(p)]
            self._spare = sp
        return sp[0] if p.data_ptr() != sp[0].data_ptr() else sp[1]

    def displacement(self, p, p_ref, relative=False):
        """max_v |p[v] - p_ref[v]| on the device (the DistMesh `ttol` test); returns a float.
        relative=True: every vertex's displacement is divided by the local mesh size fh(p[v]), which is
        what makes the test meaningful on graded meshes (lowered fh only)."""
        f = self.size.struct() if relative and self.size.lowered else None
        check(lib.dm_stage_displacement(C.byref(self.plan.c), D.ptr(p), D.ptr(p_ref), C.byref(f) if f is not None else None,
                                        D.stream_ptr()), "displacement")
        return float(self.plan.scalars()[5].item())

    def maxdp(self):
        return float(self.plan.scalars()[4].item())

    # -- pieces used at termination / by tests ----------------------------------------------
    def kept_cells(self, p, t, geps=None):
        """t[fd(centroid) < -geps] in the original order (device compaction)."""
        T = t.shape[0]
        self.ensure_plan(p.shape[0], T)
        st = D.stream_ptr()
        lv0 = self.levels[0]
        g = self.geps if geps is None else float(geps)
        keep = self.plan.keep()[:T]
        if lv0.lowered:
            check(lib.dm_cull_cells(D.ptr(lv0.prog), D.ptr(p), D.ptr(t), T, self.dim, g, D.ptr(keep), st), "cull_cells")
        else:
            cen = torch.empty((T, self.dim), dtype=torch.float64, device=p.device)
            check(lib.dm_centroids(D.ptr(p), D.ptr(t), T, self.dim, D.ptr(cen), st), "centroids")
            keep.copy_(torch.from_numpy((np.asarray(lv0.func(cen.cpu().numpy())) < -g).astype(np.uint8)))
        return compact_cells(t, keep, self.dim)

    def bars(self):
        """(E,2) int32 unique bars of the last iteration, in the reference's order."""
        check(lib.dm_stage_bar_index(C.byref(self.plan.c), D.stream_ptr()), "bar_index")
        E = self.plan.num_bars()
        pairs = torch.empty((E, 2), dtype=torch.int32, device=D.device())
        if E > 0:
            check(lib.dm_bars_pairs(C.byref(self.plan.c), D.ptr(pairs), D.stream_ptr()), "bars_pairs")
        return pairs


def bar_sizes(loop):
    """h of every unique bar of the last iteration, in bar order (tests / diagnostics)."""
    pl = loop.plan
    check(lib.dm_stage_bar_index(C.byref(pl.c), D.stream_ptr()), "bar_index")
    E = pl.num_bars()
    out = torch.empty(E, dtype=torch.float64, device=D.device())
    f = loop.size.struct()
    if E > 0:
        check(lib.dm_bar_sizes(C.byref(pl.c), C.byref(f), D.ptr(out), D.stream_ptr()), "bar_sizes")
    return out


def compact_cells(t, keep, dim):
    T = t.shape[0]
    nbytes = lib.dm_compact_scratch_bytes(T)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=t.device)
    t_out = torch.empty_like(t)
    cnt = torch.zeros(1, dtype=torch.int32, device=t.device)
    check(
        lib.dm_compact_cells(D.ptr(t), D.ptr(keep), T, dim, D.ptr(t_out), D.ptr(cnt), D.ptr(scratch), nbytes, D.stream_ptr()),
        "compact_cells",
    )
    return t_out[: int(cnt.item())]


def unique_bars(t, N=None):
    """Device replacement of `_get_edges(t)` -> `unique_edges` (mesh_generator.py:680-688):
    t (T,dim+1) int -> (E,2) int32 sorted unique (min,max) pairs.  NumPy in -> NumPy out."""
    as_torch = isinstance(t, torch.Tensor)
    td = D.to_dev(t, torch.int32)
    dim = td.shape[1] - 1
    if N is None:
        N = int(td.max().item()) + 1 if td.numel() else 1
    pl = D.Plan(N, td.shape[0], dim)
    st = D.stream_ptr()
    check(lib.dm_stage_cull_count(C.byref(pl.c), None, None, D.ptr(td), 0.0, 0, st), "cull_count")
    check(lib.dm_stage_build_adjacency(C.byref(pl.c), st), "build_bars")
    check(lib.dm_stage_bar_index(C.byref(pl.c), st), "bar_index")
    E = pl.num_bars()
    pairs = torch.empty((E, 2), dtype=torch.int32, device=td.device)
    if E > 0:
        check(lib.dm_bars_pairs(C.byref(pl.c), D.ptr(pairs), st), "bars_pairs")
    return pairs if as_torch else pairs.cpu().numpy()


def laplacian_smooth(p, t, rtol=1e-13, max_iter=200000):
    """geometry.laplacian2_fixed_point (geometry/utils.py:494-547) on the device: every interior vertex of the
    2-D mesh (p, t) at the average of its neighbours, boundary vertices fixed -- neighbour rows from stage B,
    then matrix-free preconditioned conjugate gradients (dm_laplacian_smooth).  Host arrays in and out;
    `laplacian_smooth.last` = (iterations, relative residuals)."""
    D.require_cuda()
    p = np.ascontiguousarray(p, dtype=np.float64)
    t32 = np.ascontiguousarray(t, dtype=np.int32)
    if p.ndim != 2 or p.shape[1] != 2:
        raise NotImplementedError("Laplacian smoothing only works in 2D for now")
    N, T = len(p), len(t32)
    if N == 0 or T == 0:
        return p, t
    pd, td = D.to_dev(p, torch.float64), D.to_dev(t32, torch.int32)
    plan = D.Plan(N, T, 2)
    st = D.stream_ptr()
    check(lib.dm_stage_cull_count(C.byref(plan.c), None, D.ptr(pd), D.ptr(td), 0.0, 0, st), "cull_count")  # every cell kept
    check(lib.dm_stage_build_adjacency(C.byref(plan.c), st), "build_adjacency")
    nbytes = lib.dm_laplacian_work_bytes(N)
    work = torch.empty(nbytes + 256, dtype=torch.uint8, device=D.device())
    wptr = (work.data_ptr() + 255) & ~255
    iters, resid = C.c_int(0), (C.c_double * 2)()
    check(lib.dm_laplacian_smooth(C.byref(plan.c), D.ptr(td), T, D.ptr(pd), C.c_void_p(wptr), nbytes, float(rtol), int(max_iter),
                                  C.byref(iters), resid, st), "dm_laplacian_smooth")
    laplacian_smooth.last = (iters.value, (resid[0], resid[1]))
    return pd.cpu().numpy(), t
