"""Multi-GPU: 1-D slab decomposition with a per-iteration halo exchange.

The reference's only parallel strategy is a slab decomposition along ``axis`` with ghost-vertex
export to rank +-1 over blocking pickled MPI messages (decomp/blocker.py:4-111,
migration/migration.py:116-183, mesh_generator.py:715-731,867-877).  Here: one process per GPU
(torch.distributed, NCCL over NVLink), every rank owns the vertices of one slab plus ghost copies
of its neighbours' vertices within ``width`` of the slab face; after each force iteration the
owners send the new coordinates of their exported vertices to rank +-1 with ONE grouped
``ncclSend/ncclRecv`` (``batch_isend_irecv``) and the receivers overwrite their ghost rows.
Payloads are halo coordinates only (tens of KB - MB): latency bound, so there is no data-path
collective besides this neighbour exchange.

All index bookkeeping (who owns / exports / ghosts what) is host NumPy done once per
triangulation; pack / unpack are row gathers on the device.
"""
import numpy as np
import torch
import torch.distributed as dist

__all__ = ["slab_bounds", "slab_partition", "RingHalo", "PeerHalo", "DirectHalo", "SlabLayout", "TorchComm", "generate_mesh_parallel"]


def slab_bounds(lo, hi, world):
    """Equal-width slab faces along the decomposition axis (generation/utils.py:34-42)."""
    return np.linspace(lo, hi, world + 1)


class SlabLayout:
    """Index sets of one rank.  Local vertex order: [owned | ghosts from below | ghosts from above]."""

    def __init__(self, owned, ghost_below, ghost_above, export_below, export_above):
        self.owned = owned                  # global ids of owned vertices (ascending)
        self.ghost_below = ghost_below      # global ids owned by rank-1 that are ghosts here
        self.ghost_above = ghost_above
        self.export_below = export_below    # LOCAL rows (into owned) sent to rank-1
        self.export_above = export_above    # LOCAL rows sent to rank+1

    @property
    def local_ids(self):
        return np.concatenate([self.owned, self.ghost_below, self.ghost_above])

    @property
    def n_owned(self):
        return len(self.owned)


def slab_partition(coord, faces, rank, width):
    """Partition by the coordinate along the decomposition axis.

    coord : (N_global,) coordinate of every global vertex along `axis`
    faces : (world+1,) slab faces; rank r owns faces[r] <= x < faces[r+1] (last slab closed)
    width : halo width (the reference pads extents by 5*h0, mesh_generator.py:873-874)

    A vertex owned by rank r is exported to r-1 if x < faces[r] + width and to r+1 if
    x >= faces[r+1] - width; the neighbour's ghost list is the same set, in ascending global id, so
    both sides agree on the message layout without any handshake.
    """
    world = len(faces) - 1
    own_of = np.clip(np.searchsorted(faces, coord, side="right") - 1, 0, world - 1)
    ids = np.arange(len(coord))
    owned = ids[own_of == rank]
    x = coord[owned]
    exp_b = np.nonzero(x < faces[rank] + width)[0] if rank > 0 else np.zeros(0, dtype=np.int64)
    exp_a = np.nonzero(x >= faces[rank + 1] - width)[0] if rank < world - 1 else np.zeros(0, dtype=np.int64)
    if rank > 0:
        below = ids[own_of == rank - 1]
        ghost_b = below[coord[below] >= faces[rank] - width]
    else:
        ghost_b = np.zeros(0, dtype=np.int64)
    if rank < world - 1:
        above = ids[own_of == rank + 1]
        ghost_a = above[coord[above] < faces[rank + 1] + width]
    else:
        ghost_a = np.zeros(0, dtype=np.int64)
    return SlabLayout(owned, ghost_b, ghost_a, exp_b, exp_a)


class RingHalo:
    """Per-iteration neighbour exchange of ghost coordinates on a 1-D chain of ranks."""

    def __init__(self, layout, dim, device, rank=None, world=None, group=None):
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.group = group
        self.dim = dim
        n0 = layout.n_owned
        nb, na = len(layout.ghost_below), len(layout.ghost_above)
        self.exp_b = torch.as_tensor(layout.export_below, dtype=torch.int64, device=device)
        self.exp_a = torch.as_tensor(layout.export_above, dtype=torch.int64, device=device)
        self.ghost_b = slice(n0, n0 + nb)
        self.ghost_a = slice(n0 + nb, n0 + nb + na)
        self.send_b = torch.empty((len(self.exp_b), dim), dtype=torch.float64, device=device)
        self.send_a = torch.empty((len(self.exp_a), dim), dtype=torch.float64, device=device)
        self.recv_b = torch.empty((nb, dim), dtype=torch.float64, device=device)
        self.recv_a = torch.empty((na, dim), dtype=torch.float64, device=device)
        self.bytes_per_exchange = 8 * dim * (len(self.exp_b) + len(self.exp_a) + nb + na)

    def exchange(self, p):
        """Send the rows of `p` exported to rank+-1, overwrite the ghost rows with what arrives."""
        ops = []
        if self.rank > 0:
            torch.index_select(p, 0, self.exp_b, out=self.send_b)
            ops.append(dist.P2POp(dist.isend, self.send_b, self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv_b, self.rank - 1, self.group))
        if self.rank < self.world - 1:
            torch.index_select(p, 0, self.exp_a, out=self.send_a)
            ops.append(dist.P2POp(dist.isend, self.send_a, self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv_a, self.rank + 1, self.group))
        if not ops:
            return
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if self.rank > 0:
            p[self.ghost_b] = self.recv_b
        if self.rank < self.world - 1:
            p[self.ghost_a] = self.recv_a


class PeerHalo:
    """The same neighbour exchange as :class:`RingHalo`, over NVLink peer memory instead of NCCL
    messages: every rank owns a symmetric-memory ghost buffer (two slots, alternating per call);
    ``exchange`` pushes the exported rows of ``p`` straight into the neighbours' buffers with ONE
    hand-written kernel per neighbour (`dm_halo_push`: remote stores), raises a signal on each
    neighbour, waits for theirs and copies what arrived into its ghost rows.  A rank can be at most
    one call ahead of a neighbour (it waits for the neighbour's signal of the same call), so two
    slots suffice without an acknowledgement."""

    def __init__(self, layout, dim, device, rank=None, world=None, group=None):
        import ctypes as C  # noqa: F401
        import torch.distributed._symmetric_memory as symm

        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.dim = dim
        n0 = layout.n_owned
        nb, na = len(layout.ghost_below), len(layout.ghost_above)
        self.exp_b = torch.as_tensor(layout.export_below, dtype=torch.int32, device=device)
        self.exp_a = torch.as_tensor(layout.export_above, dtype=torch.int32, device=device)
        self.ghost_b = slice(n0, n0 + nb)
        self.ghost_a = slice(n0 + nb, n0 + nb + na)
        self.nb, self.na = nb, na
        cap = torch.tensor([max(nb, na, 1)], dtype=torch.int64, device=device)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)
        self.cap = int(cap.item())  # rows per side, the same on every rank (symmetric allocation)
        # [slot][side: 0 = from below, 1 = from above][cap][dim]
        self.buf = symm.empty((2, 2, self.cap, dim), dtype=torch.float64, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(a) for a in self.hdl.buffer_ptrs]
        self.slot = 0
        self.bytes_per_exchange = 8 * dim * (len(self.exp_b) + len(self.exp_a) + nb + na)
        self.hdl.barrier(channel=2)

    def _peer(self, peer, slot, side):
        return self.ptrs[peer] + 8 * self.dim * self.cap * (2 * slot + side)

    def exchange(self, p):
        import ctypes as C

        from . import device as D
        from ._lib import check, lib

        s, self.slot = self.slot, self.slot ^ 1
        st = D.stream_ptr()
        r, w = self.rank, self.world
        if r > 0:      # my "below" exports are rank r-1's ghosts "from above"
            check(lib.dm_halo_push(D.ptr(p), D.ptr(self.exp_b), len(self.exp_b), self.dim,
                                   C.c_void_p(self._peer(r - 1, s, 1)), st), "dm_halo_push")
        if r < w - 1:  # my "above" exports are rank r+1's ghosts "from below"
            check(lib.dm_halo_push(D.ptr(p), D.ptr(self.exp_a), len(self.exp_a), self.dim,
                                   C.c_void_p(self._peer(r + 1, s, 0)), st), "dm_halo_push")
        if r > 0:
            self.hdl.put_signal(r - 1, channel=1)
        if r < w - 1:
            self.hdl.put_signal(r + 1, channel=0)
        if r > 0:
            self.hdl.wait_signal(r - 1, channel=0)
            p[self.ghost_b] = self.buf[s, 0, : self.nb]
        if r < w - 1:
            self.hdl.wait_signal(r + 1, channel=1)
            p[self.ghost_a] = self.buf[s, 1, : self.na]


class DirectHalo:
    """The neighbour exchange as TWO launches per iteration, with no staging buffer and no unpack copy:
    every rank keeps its position buffers in symmetric memory (`slot(k)`: where the force iteration of
    step k writes its result); `exchange(p)` stores the rows this rank exports straight into the GHOST
    ROWS of the neighbours' buffers of the same step (`dm_halo_push2`: NVLink stores + a stamp raised on
    each neighbour by the kernel's last block) and then holds the stream until both neighbours' stamps
    of this step have arrived (`dm_halo_wait`).  Two alternating slots suffice: a neighbour can be at most
    one step ahead (it waits for this rank's stamp before its next iteration), and then it writes ghost rows
    of the slot this rank is still computing its OWNED rows of.  Same values as RingHalo / PeerHalo."""

    def __init__(self, layout, dim, device, rank=None, world=None, group=None):
        import torch.distributed._symmetric_memory as symm

        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.dim = dim
        n0, nb, na = layout.n_owned, len(layout.ghost_below), len(layout.ghost_above)
        self.n_local = n0 + nb + na
        self.exp_b = torch.as_tensor(layout.export_below, dtype=torch.int32, device=device)
        self.exp_a = torch.as_tensor(layout.export_above, dtype=torch.int32, device=device)
        counts = torch.zeros((self.world, 3), dtype=torch.int64, device=device)
        counts[self.rank] = torch.tensor([n0, nb, na], dtype=torch.int64, device=device)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        self.counts = counts.cpu().numpy()
        self.cap = int(self.counts.sum(axis=1).max())  # rows per slot, the same on every rank (symmetric allocation)
        grp = group if group is not None else dist.group.WORLD
        self.buf = symm.empty((2, self.cap, dim), dtype=torch.float64, device=device)
        self.flags = symm.empty((2,), dtype=torch.int64, device=device)  # stamps: [from below, from above]
        self.flags.zero_()
        self.hbuf = symm.rendezvous(self.buf, grp)
        self.hflag = symm.rendezvous(self.flags, grp)
        self.pbuf = [int(a) for a in self.hbuf.buffer_ptrs]
        self.pflag = [int(a) for a in self.hflag.buffer_ptrs]
        self.done = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.step = 0
        self.bytes_per_exchange = 8 * dim * (len(self.exp_b) + len(self.exp_a) + nb + na)
        torch.cuda.synchronize()
        self.hbuf.barrier(channel=3)

    def slot(self, k=None):
        """(n_local, dim) position buffer of step k (default: the step the next exchange belongs to)."""
        k = self.step if k is None else k
        return self.buf[k % 2, : self.n_local]

    def exchange(self, p):
        """Push the exported rows of `p` into the neighbours' slot of this step and wait for theirs; the ghost
        rows of `self.slot()` then hold the neighbours' new positions."""
        import ctypes as C

        from . import device as D
        from ._lib import check, lib

        r, w, s = self.rank, self.world, self.step % 2
        row = 8 * self.dim
        stamp = self.step + 1
        db = fb = da = fa = None
        if r > 0:      # my "below" exports are rank r-1's ghosts "from above": behind its owned + from-below rows
            n0p, nbp, _ = self.counts[r - 1]
            db = C.c_void_p(self.pbuf[r - 1] + row * (s * self.cap + int(n0p + nbp)))
            fb = C.c_void_p(self.pflag[r - 1] + 8)
        if r < w - 1:  # my "above" exports are rank r+1's ghosts "from below": right behind its owned rows
            n0p = self.counts[r + 1][0]
            da = C.c_void_p(self.pbuf[r + 1] + row * (s * self.cap + int(n0p)))
            fa = C.c_void_p(self.pflag[r + 1])
        st = D.stream_ptr()
        check(lib.dm_halo_push2(D.ptr(p), self.dim, D.ptr(self.exp_b), len(self.exp_b) if r > 0 else 0, db, fb,
                                D.ptr(self.exp_a), len(self.exp_a) if r < w - 1 else 0, da, fa, stamp, D.ptr(self.done), st),
              "dm_halo_push2")
        mine = self.flags.data_ptr()
        check(lib.dm_halo_wait(C.c_void_p(mine) if r > 0 else None, C.c_void_p(mine + 8) if r < w - 1 else None, stamp,
                               D.ptr(self.err), st), "dm_halo_wait")
        self.step += 1

    def check(self):
        if int(self.err.item()) != 0:
            raise RuntimeError("DirectHalo: a neighbour's stamp did not arrive (dm_halo_wait timed out)")


def make_slab_workload(workload, h0, rank, world, halo_layers=5, seed=0):
    """Synthetic weak-scaling input for bench.py --gpus N: `world` slabs along axis 1 of a
    cylinder (3-D, axis = y) / rectangle (2-D), each slab with the volume (area) of the unit ball
    (disk), i.e. the per-GPU work of the N=1 workload.  Returns (local points [owned|ghosts],
    dim, domain object, SlabLayout).  The reference pads slab extents by 5*h0
    (mesh_generator.py:873-874), hence halo_layers=5.  Every rank builds the same global jittered
    lattice (seeded), so owners and ghost holders agree on positions without communication."""
    from . import geometry

    if workload == "ball":
        dim, ell = 3, 4.0 / 3.0
    else:
        dim, ell = 2, np.pi / 2.0
    Ly = world * ell
    lo = [-1.0] * dim
    hi = [1.0] * dim
    lo[1], hi[1] = -Ly / 2, Ly / 2
    axes = [np.arange(int(np.ceil((b + h0 - a) / h0)), dtype=float) * h0 + a for a, b in zip(lo, hi)]
    g = [a.copy() for a in np.meshgrid(*axes, indexing="ij")]
    g[1][1::2] += h0 / 2
    if dim == 3:
        g[2][1::2] += h0 / 2
    p = np.stack([a.ravel() for a in g], axis=1)
    if dim == 3:
        dom = geometry.Cylinder(h=Ly, r=1.0)
        rad = np.sqrt(p[:, 0] ** 2 + p[:, 2] ** 2)
        inside = (rad - 1.0 < 0.1 * h0) & (np.abs(p[:, 1]) - Ly / 2 < 0.1 * h0)
    else:
        dom = geometry.Rectangle((-1.0, 1.0, -Ly / 2, Ly / 2))
        inside = (np.abs(p[:, 0]) - 1.0 < 0.1 * h0) & (np.abs(p[:, 1]) - Ly / 2 < 0.1 * h0)
    p = p[inside]
    rng = np.random.default_rng(seed)
    p = p + rng.uniform(-0.1 * h0, 0.1 * h0, p.shape)
    faces = slab_bounds(-Ly / 2 - h0, Ly / 2 + h0, world)
    faces[1:-1] = np.linspace(-Ly / 2, Ly / 2, world + 1)[1:-1]
    layout = slab_partition(p[:, 1], faces, rank, width=halo_layers * h0)
    return np.ascontiguousarray(p[layout.local_ids]), dim, dom, layout


# ----------------------------------------------------------------------------------------------
# generate_mesh on several GPUs: the reference's parallel algorithm (mesh_generator.py:430-530,
# 715-731, 808-880; migration/migration.py:72-183) with one process per GPU over torch.distributed.
# ----------------------------------------------------------------------------------------------
class TorchComm:
    """What to pass as ``comm=`` : the mpi4py-like (rank, size) view of a torch.distributed group."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)


def _comm_device(dev):
    """Tensors that travel live on the GPU under NCCL and on the host under gloo (CPU tests)."""
    return dev if dist.get_backend() == "nccl" else torch.device("cpu")


def _exchange_ghosts(below, above, rank, size, dim, cdev, group=None):
    """Send `below` / `above` (host arrays (k,dim)) to rank-1 / rank+1, return what arrives as
    (from_above, from_below) -- the message pattern of migration.exchange (migration.py:148-183):
    counts first, then ONE grouped send/recv of the coordinates."""
    nb = torch.tensor([len(below)], dtype=torch.int64, device=cdev)
    na = torch.tensor([len(above)], dtype=torch.int64, device=cdev)
    rb = torch.zeros(1, dtype=torch.int64, device=cdev)
    ra = torch.zeros(1, dtype=torch.int64, device=cdev)
    ops = []
    if rank > 0:
        ops += [dist.P2POp(dist.isend, nb, rank - 1, group), dist.P2POp(dist.irecv, rb, rank - 1, group)]
    if rank < size - 1:
        ops += [dist.P2POp(dist.isend, na, rank + 1, group), dist.P2POp(dist.irecv, ra, rank + 1, group)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    sb = torch.from_numpy(np.ascontiguousarray(below, dtype=np.float64)).to(cdev)
    sa = torch.from_numpy(np.ascontiguousarray(above, dtype=np.float64)).to(cdev)
    gb = torch.empty((int(rb.item()), dim), dtype=torch.float64, device=cdev)
    ga = torch.empty((int(ra.item()), dim), dtype=torch.float64, device=cdev)
    ops = []
    if rank > 0:
        if sb.numel():
            ops.append(dist.P2POp(dist.isend, sb, rank - 1, group))
        if gb.numel():
            ops.append(dist.P2POp(dist.irecv, gb, rank - 1, group))
    if rank < size - 1:
        if sa.numel():
            ops.append(dist.P2POp(dist.isend, sa, rank + 1, group))
        if ga.numel():
            ops.append(dist.P2POp(dist.irecv, ga, rank + 1, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return ga.cpu().numpy(), gb.cpu().numpy()


def _broadcast_points(points, rank, dim, cdev, group=None):
    """rank 0's (N,dim) array on every rank (the reference reads `points` on rank 0 only and sends each
    rank its block, migration.localize_points, migration.py:55-69)."""
    n = torch.tensor([0 if points is None else len(points)], dtype=torch.int64, device=cdev)
    dist.broadcast(n, 0, group=group)
    if rank == 0:
        buf = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64).reshape(-1, dim)).to(cdev)
    else:
        buf = torch.empty((int(n.item()), dim), dtype=torch.float64, device=cdev)
    dist.broadcast(buf, 0, group=group)
    return buf.cpu().numpy()


def _gather_points(points, rank, size, dim, cdev, group=None):
    """Owned vertices of every rank -> one array on rank 0, in rank order (None elsewhere)."""
    counts = torch.zeros(size, dtype=torch.int64, device=cdev)
    counts[rank] = len(points)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    counts = counts.cpu().numpy()
    if rank != 0:
        buf = torch.from_numpy(np.ascontiguousarray(points)).to(cdev)
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, buf, 0, group)]):
            w.wait()
        return None
    bufs = [torch.empty((int(counts[r]), dim), dtype=torch.float64, device=cdev) for r in range(1, size)]
    for w in dist.batch_isend_irecv([dist.P2POp(dist.irecv, b, r + 1, group) for r, b in enumerate(bufs)]):
        w.wait()
    return np.ascontiguousarray(np.vstack([points] + [b.cpu().numpy() for b in bufs]))


def generate_mesh_parallel(domain, edge_length, comm, **kwargs):  # noqa: C901
    """generate_mesh over `comm.size` ranks, one GPU each: every rank owns the vertices created in
    its slab of the bounding box (equal-width slabs along ``axis``, generation/utils.py:28-44); per
    iteration it triangulates its vertices, exports those whose incident-cell circumballs reach a
    neighbour's (5*h0-padded) extent (`dm_halo_select`, replacing cpputils.where_to2/3), receives its
    neighbours' exports as ghosts, retriangulates owned+ghosts, drops the cells that lie entirely
    outside its extent, runs the force iteration on the device and keeps the owned rows.  At
    `max_iter` rank 0 gathers the owned vertices, triangulates them once and cleans up; the other ranks
    return (True, True) (mesh_generator.py:460-527, migration.py:72-183).
    The mesh-size grid is replicated on every GPU instead of being resampled per slab
    (migration.localize_sizing_function), so parallel and serial runs see the same fh."""
    import ctypes as C
    import time
    import warnings

    from . import device as D
    from . import generation as G
    from . import geometry
    from ._lib import check, lib
    from .engine import ForceLoop, Level
    from .triangulator import _Rebuilt, get_triangulator, host_threads

    rank, size_ = int(comm.rank), int(comm.size)
    group = getattr(comm, "group", None)
    if not dist.is_initialized() or dist.get_world_size(group) != size_:
        raise RuntimeError("generate_mesh(comm=...) with comm.size > 1 needs torch.distributed initialised with that many ranks")
    gen_opts = {
        "verbose": 1, "max_iter": 50, "seed": 0, "perform_checks": False, "pfix": None, "axis": 1,
        "points": None, "delta_t": 0.30, "geps_mult": 0.1, "subdomains": None, "mesh_improvement": True,
        "r0m_is_h0": False, "triangulator": None, "ttol": None,
    }
    return_state = bool(kwargs.pop("_return_state", False))  # bench.py: stop before the gather, hand the slab over
    gen_opts.update(kwargs)
    G._parse_kwargs(kwargs)
    if gen_opts["pfix"] is not None and rank == 0:
        # the reference drops fixed points when comm.size > 1 (_unpack_pfix, mesh_generator.py:880-888)
        warnings.warn("`pfix` is ignored when comm.size > 1 (as in the reference)")
    if gen_opts["ttol"] is not None and rank == 0:
        warnings.warn("`ttol` is ignored when comm.size > 1: every iteration retriangulates (twice, with the ghosts)")
    print_msg1, print_msg2 = G._printers(gen_opts)
    if rank != 0:
        print_msg1 = print_msg2 = lambda msg: None  # noqa: E731

    dom, bbox0, _ = G._unpack_domain(domain, gen_opts)
    payload, bbox1, hmin = G._unpack_sizing(edge_length)
    bbox = bbox0 if bbox1 is None else G._minmax(bbox0, bbox1)
    if not isinstance(bbox, tuple):
        raise ValueError("`bbox` must be a tuple")
    dim = int(len(bbox) / 2)
    if bbox0 != bbox1 and bbox1 is not None:
        dom = geometry.Rectangle(bbox) if dim == 2 else geometry.Cube(bbox)
    bbox_arr = np.array(bbox, dtype=float).reshape(-1, 2)
    h0 = hmin if hmin is not None else gen_opts["h0"]
    if h0 < 0:
        raise ValueError("`h0` must be > 0")
    if gen_opts["max_iter"] < 0:
        raise ValueError("`max_iter` must be > 0")
    max_iter, axis = gen_opts["max_iter"], gen_opts["axis"]
    geps = gen_opts["geps_mult"] * h0
    deps = np.sqrt(np.finfo(np.double).eps) * h0
    level0 = Level(dom, dim)
    size = G._size_spec(payload, dim)
    levels = [level0]
    if gen_opts["subdomains"] is not None:
        for sub in gen_opts["subdomains"]:
            levels.append(Level(G._unpack_domain(sub, gen_opts)[0], dim))
    dev = D.device()
    cdev = _comm_device(dev)

    def all_ranks_ok(ok, what):
        """A failed check must stop EVERY rank (a rank that raises alone leaves the others blocked in
        the next collective): agree on the outcome first, then raise everywhere."""
        flag = torch.tensor([0 if ok else 1], dtype=torch.int64, device=cdev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        assert int(flag.item()) == 0, what

    if gen_opts["points"] is None:
        # ---- initial points of this rank's slab (make_init_points + rejection, :808-852) ----
        lb = bbox_arr.copy()
        lims = np.linspace(lb[axis, 0], lb[axis, 1], size_ + 1)
        lb[axis, :] = lims[rank : rank + 2]
        if rank != 0:  # "starting point must be lasts + h0"
            prev = lims[rank - 1 : rank + 1]
            lb[axis, 0] = prev[0] + (int(np.ceil((prev[1] + h0 - prev[0]) / h0)) - 1) * h0 + h0
        p = G._staggered_grid(h0, dim, lb)
        p = p[level0.eval_host(p) < geps]
        r0 = size.eval_host(p)
        r0m = float(r0.min()) if len(r0) else np.inf
        if gen_opts["r0m_is_h0"]:
            r0m = 1.1 * h0 if 1.1 * h0 < r0m else h0
        t_r0m = torch.tensor([r0m], dtype=torch.float64, device=cdev)
        dist.all_reduce(t_r0m, op=dist.ReduceOp.MIN, group=group)  # "decimation occurs uniformly across ranks"
        r0m = float(t_r0m.item())
        np.random.seed(gen_opts["seed"])
        p = np.ascontiguousarray(p[np.random.rand(p.shape[0]) < r0m**dim / r0**dim])
        pad = 5 * h0  # _form_extents pads the AABB of the owned points by 5*h0 along the axis (:867-877)
        ext_axis = axis
    else:
        # ---- user-defined points (restart; _user_defined_points :787-805): rank 0's array is cut into
        # comm.size blocks (decomp.blocker, decomp/blocker.py:4-111: `axis` 1 cuts along x, 0 along y,
        # 2 along z; equal-width blocks of the points' bounding box) and every rank takes its block.
        # blocker returns the blocks' bounding boxes UNPADDED, and the export test (here and in the reference,
        # cpputils.cpp:122-135) only looks at cells whose bounding box overlaps the neighbour's extent: with
        # disjoint extents nothing would ever be exported.  The extents are therefore padded by 5*h0 along the
        # cut, exactly as _form_extents does for generated points (:867-877).
        pts = _broadcast_points(gen_opts["points"], rank, dim, cdev, group)
        ext_axis = {0: 1, 1: 0, 2: 2}[axis]
        if ext_axis >= dim:
            raise ValueError("Dimensions of points are not supported")
        eps_ = np.finfo(float).eps
        lo_, hi_ = pts[:, ext_axis].min() - eps_, pts[:, ext_axis].max() + eps_
        cuts = np.linspace(lo_, hi_, size_ + 1)
        blk = np.clip(np.searchsorted(cuts, pts[:, ext_axis], side="right") - 1, 0, size_ - 1)
        p = np.ascontiguousarray(pts[blk == rank])
        pad = 5 * h0
    all_ranks_ok(len(p) > 0, "No vertices to mesh with!")
    # extents of every rank: bounding box of its points (padded, see above)
    ext = torch.zeros((size_, 2 * dim), dtype=torch.float64, device=cdev)
    mine = np.concatenate([p.min(0), p.max(0)])
    mine[ext_axis] -= pad
    mine[ext_axis + dim] += pad
    ext[rank] = torch.from_numpy(mine).to(cdev)
    dist.all_reduce(ext, op=dist.ReduceOp.SUM, group=group)
    extents = ext.cpu().numpy()
    print_msg1("Commencing mesh generation with %d vertices on rank %d." % (len(p), rank))

    tri = get_triangulator(gen_opts["triangulator"], dim)
    loop = ForceLoop(dim, levels, size, h0, geps, deps, delta_t=gen_opts["delta_t"], nfix=0)
    boxes = np.zeros((2, 2 * dim))
    if rank > 0:
        boxes[0] = extents[rank - 1]
    if rank < size_ - 1:
        boxes[1] = extents[rank + 1]
    boxes_c = (C.c_double * (4 * dim))(*boxes.ravel())
    stats = dict(delaunay=0.0, device=0.0, exchange=0.0, iterations=0, nverts=len(p), triangulator=tri.name)
    count = 0
    while True:
        start = time.time()
        n_own = len(p)
        t0 = time.perf_counter()
        dt = tri.build(p) if hasattr(tri, "build") else _Rebuilt(tri, p)  # (stays around: the ghosts are inserted into it below, as the reference does with its CGAL object)
        t_own = dt.cells()
        stats["delaunay"] += time.perf_counter() - t0
        # ---- ghosts: export the vertices whose cells' circumballs reach a neighbour (enqueue + exchange)
        t1 = time.perf_counter()
        pd = D.to_dev(p, torch.float64)
        td = D.to_dev(t_own, torch.int32)
        flags = torch.zeros((n_own + 3) // 4 * 4, dtype=torch.uint8, device=dev)
        check(lib.dm_halo_select(D.ptr(pd), D.ptr(td), td.shape[0], n_own, dim, boxes_c, int(rank > 0),
                                 int(rank < size_ - 1), D.ptr(flags), D.stream_ptr()), "dm_halo_select")
        fl = flags[:n_own].cpu().numpy()
        from_above, from_below = _exchange_ghosts(p[(fl & 1) != 0], p[(fl & 2) != 0], rank, size_, dim, cdev, group)
        stats["exchange"] += time.perf_counter() - t1
        # local vertex order: [owned | ghosts from below | ghosts from above]
        ghosts = [g for g in (from_below, from_above) if len(g)]
        p_loc = np.ascontiguousarray(np.vstack([p] + ghosts)) if ghosts else p
        t0 = time.perf_counter()
        if ghosts:
            for g in ghosts:
                dt.insert(g)
            t_loc = dt.cells()
        else:
            t_loc = t_own
        dt.close()
        stats["delaunay"] += time.perf_counter() - t0
        # cells with all their vertices outside this rank's extent belong to somebody else
        # (geometry.remove_external_entities, geometry/utils.py:57-94)
        e = extents[rank]
        outside = ((p_loc < e[:dim]) | (p_loc > e[dim:])).any(axis=1)
        t_loc = np.ascontiguousarray(t_loc[~outside[t_loc].all(axis=1)])
        pd = D.to_dev(p_loc, torch.float64)
        td = D.to_dev(t_loc, torch.int32)

        if count == (max_iter - 1) and return_state:
            # the slab as the next force iteration would see it (bench.py's multi-GPU step): local points
            # and cells, what this rank exports to rank-1 / rank+1 (rows of the owned block, in the order
            # the neighbours hold them as ghosts) and how many ghosts it holds from each side
            G.last_run_stats.clear()
            G.last_run_stats.update(stats)
            return dict(p=p_loc, t=t_loc, n_owned=n_own, export_below=np.nonzero(fl & 1)[0], export_above=np.nonzero(fl & 2)[0],
                        n_ghost_below=len(from_below), n_ghost_above=len(from_above), extents=extents, dim=dim, h0=h0)
        if count == (max_iter - 1):
            # The reference gathers the local meshes (ghost copies included) and de-duplicates them on
            # rank 0 (migration.aggregate + fix_mesh); here rank 0 gathers the OWNED vertices and
            # triangulates them once: the same vertex set, and a conforming mesh by construction.
            print_msg1("Termination reached...maximum number of iterations reached.")
            gp = _gather_points(p, rank, size_, dim, cdev, group)
            G.last_run_stats.clear()
            G.last_run_stats.update(stats)
            if rank != 0:
                return True, True
            # (the other ranks are done: the final triangulation may use the threads they no longer need)
            whole = get_triangulator(gen_opts["triangulator"], dim, threads=host_threads(ranks=1)) if not hasattr(gen_opts["triangulator"], "triangulate") else tri
            gt = whole.triangulate(gp)
            t_kept = loop.kept_cells(D.to_dev(gp, torch.float64), D.to_dev(gt, torch.int32)).cpu().numpy()
            p_out, t_out = G._termination(gp, t_kept, gen_opts, dim, verbose=gen_opts["verbose"])
            p_out = G._level_set_newton(p_out, t_out, level0, deps, dim)
            fin = ForceLoop(dim, [level0], size, h0, h0 * 0.001, deps)
            t_out = fin.kept_cells(D.to_dev(p_out, torch.float64), D.to_dev(t_out, torch.int32)).cpu().numpy().astype(t_out.dtype)
            return p_out, t_out

        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        loop.n_rows = n_own  # rows, forces and the update for the owned vertices only; ghosts are just neighbours
        p_new, _ = loop.iterate(pd, td)
        ev1.record()
        p = np.ascontiguousarray(p_new[:n_own].cpu().numpy())  # "delete ghost points"
        stats["device"] += ev0.elapsed_time(ev1) * 1e-3
        stats["iterations"] += 1
        maxdp = loop.maxdp()
        print_msg2("Iteration #%d, max movement is %f, there are %d vertices and %d cells" % (count + 1, maxdp, len(p_loc), len(t_loc)))
        all_ranks_ok(maxdp < 1000 * h0, "max movement indicates there's a convergence problem")
        count += 1
        print_msg2("     Elapsed wall-clock time %f : " % (time.time() - start))
