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

__all__ = ["slab_bounds", "slab_partition", "RingHalo", "SlabLayout"]


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


def allreduce_force_scale(sum_L, sum_h, group=None):
    """Optional: global (sum L^d, sum h^d) so that every slab uses the single-GPU force scale
    (the reference uses rank-local sums, mesh_generator.py:700)."""
    buf = torch.stack([sum_L, sum_h])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[0], buf[1]
