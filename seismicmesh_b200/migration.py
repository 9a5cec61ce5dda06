"""`SeismicMesh.migration` (migration/migration.py:16-183) for one process per GPU over
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests): the reference's
function names and meaning, with the ghost selection as a CUDA kernel (`dm_halo_select`, replacing
the CGAL-based `cpputils.where_to2/3`, migration/cpp/cpputils.cpp:85-200,247-383) and grouped
send/recv instead of pickled blocking MPI messages.  `comm` is a :class:`parallel.TorchComm`."""
import ctypes as C

import numpy as np
import torch

from . import parallel as _par

__all__ = ["enqueue", "aggregate", "exchange", "localize_points", "localize_sizing_function"]


def localize_sizing_function(fh, h0, bbox, dim, axis, comm):
    """The reference resamples fh onto a per-slab grid of spacing h0 and ships it to the slab's rank
    (:16-52).  A B200 holds the full grid (<= ~1 GB of 180 GB), so the size function is replicated:
    every rank keeps `fh` as it is and parallel runs see the same fh as serial ones."""
    return fh


def localize_points(blocks, extents, comm, dim):
    """Rank r receives `blocks[r]` and everybody the extents (:55-69); rank 0 holds the inputs."""
    import torch.distributed as dist

    rank, size, group = int(comm.rank), int(comm.size), getattr(comm, "group", None)
    cdev = _par._comm_device(torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
    counts = torch.zeros(size, dtype=torch.int64, device=cdev)
    ext = torch.zeros((size, 2 * dim), dtype=torch.float64, device=cdev)
    if rank == 0:
        counts[:] = torch.tensor([len(b) for b in blocks], dtype=torch.int64)
        ext[:] = torch.tensor(np.asarray(extents, dtype=np.float64))
    dist.broadcast(counts, 0, group=group)
    dist.broadcast(ext, 0, group=group)
    if rank == 0:
        ops = [dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(blocks[r], dtype=np.float64)).to(cdev), r, group)
               for r in range(1, size)]
        mine = np.ascontiguousarray(blocks[0], dtype=np.float64)
    else:
        buf = torch.empty((int(counts[rank].item()), dim), dtype=torch.float64, device=cdev)
        ops = [dist.P2POp(dist.irecv, buf, 0, group)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if rank != 0:
        mine = buf.cpu().numpy()
    return mine, [list(e) for e in ext.cpu().numpy()]


def enqueue(extents, points, faces, rank, size, dim=2):
    """Which vertices have to be sent to rank-1 / rank+1: those with an incident cell whose circumball
    touches the neighbour's extent (:116-145).  Returns the reference's packed table: row 0 =
    [NSB, NSA, 0...], then the NSB vertices for the rank below and the NSA for the rank above, one per
    row as (x, y[, z], vertex id)."""
    from . import device as D
    from ._lib import check, lib

    p = D.points_dev(points, dim)
    t = D.to_dev(np.ascontiguousarray(faces), torch.int32)
    n = p.shape[0]
    boxes = np.zeros((2, 2 * dim))
    if rank > 0:
        boxes[0] = np.asarray(extents[rank - 1], dtype=np.float64)
    if rank < size - 1:
        boxes[1] = np.asarray(extents[rank + 1], dtype=np.float64)
    flags = torch.zeros((n + 3) // 4 * 4, dtype=torch.uint8, device=p.device)
    check(lib.dm_halo_select(D.ptr(p), D.ptr(t), t.shape[0], n, dim, (C.c_double * (4 * dim))(*boxes.ravel()),
                             int(rank > 0), int(rank < size - 1), D.ptr(flags), D.stream_ptr()), "dm_halo_select")
    fl = flags[:n].cpu().numpy()
    pts = np.asarray(points, dtype=np.float64)
    below, above = np.nonzero(fl & 1)[0], np.nonzero(fl & 2)[0]
    out = np.zeros((1 + len(below) + len(above), dim + 1))
    out[0, 0], out[0, 1] = len(below), len(above)
    out[1 : 1 + len(below), :dim], out[1 : 1 + len(below), dim] = pts[below], below
    out[1 + len(below) :, :dim], out[1 + len(below) :, dim] = pts[above], above
    return out


def exchange(comm, rank, size, exports, dim=2):
    """Send the enqueued vertices to rank-1 / rank+1 and return what the neighbours sent here, the
    points from above first (:148-183): ONE grouped send/recv after the counts."""
    nsb, nsa = int(exports[0, 0]), int(exports[0, 1])
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    from_above, from_below = _par._exchange_ghosts(exports[1 : 1 + nsb, :dim], exports[1 + nsb : 1 + nsb + nsa, :dim], rank,
                                                   size, dim, _par._comm_device(dev), getattr(comm, "group", None))
    got = [g for g in (from_above, from_below) if len(g)]
    return np.vstack(got) if got else np.array([[]])


def aggregate(points, faces, comm, size, rank, dim=2):
    """Collect the local meshes on rank 0 (:72-112): every rank cleans its mesh (fix_mesh), rank 0
    receives points and cells (renumbered by the running vertex offset) in rank order; the other ranks
    return (True, True)."""
    import torch.distributed as dist

    from . import meshutil

    group = getattr(comm, "group", None)
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    cdev = _par._comm_device(dev)
    points, faces, _ = meshutil.fix_mesh(np.asarray(points, dtype=np.float64), np.asarray(faces), delete_unused=True, dim=dim)
    gp = _par._gather_points(points, rank, size, dim, cdev, group)
    gf = _par._gather_points(np.asarray(faces, dtype=np.float64), rank, size, dim + 1, cdev, group)
    counts = torch.zeros(size, dtype=torch.int64, device=cdev)
    counts[rank] = len(points)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    nf = torch.zeros(size, dtype=torch.int64, device=cdev)
    nf[rank] = len(faces)
    dist.all_reduce(nf, op=dist.ReduceOp.SUM, group=group)
    if rank != 0:
        return True, True
    off = np.concatenate([[0], np.cumsum(counts.cpu().numpy())])
    gf = gf.astype(np.int64)
    start = 0
    for r, k in enumerate(nf.cpu().numpy()):
        gf[start : start + k] += off[r]
        start += k
    return gp, gf
