"""Multi-rank host logic on CPU: slab partition + ring halo exchange with world_size 2 and 3 over
`gloo` (the N>1 path uses NCCL on the GPU box; the index bookkeeping and the message pattern are
identical)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seismicmesh_b200.parallel import RingHalo, slab_bounds, slab_partition


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _global_points(n=4000, dim=3, seed=0):
    rng = np.random.default_rng(seed)
    p = rng.random((n, dim))
    p[:, 1] *= 4.0  # long along the decomposition axis
    return p


def _worker(rank, world, port, dim, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pg = _global_points(dim=dim)
        faces = slab_bounds(0.0, 4.0, world)
        lay = slab_partition(pg[:, 1], faces, rank, width=0.3)
        ids = lay.local_ids
        p = torch.from_numpy(pg[ids].copy())
        # every rank "updates" its OWNED rows with a rank-independent function of the global id;
        # ghost rows are poisoned and must be repaired by the exchange
        new = lambda gid: np.stack([np.sin(gid * 0.1 + k) for k in range(dim)], axis=1)  # noqa: E731
        p[: lay.n_owned] = torch.from_numpy(new(lay.owned.astype(np.float64)))
        p[lay.n_owned :] = float("nan")
        halo = RingHalo(lay, dim, torch.device("cpu"), rank=rank, world=world)
        halo.exchange(p)
        expect = new(ids.astype(np.float64))
        ok = bool(np.array_equal(p.numpy(), expect))
        q.put((rank, ok, lay.n_owned, len(lay.ghost_below), len(lay.ghost_above), halo.bytes_per_exchange))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dim", [(2, 3), (3, 2)])
def test_ring_halo_exchange_gloo(world, dim):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dim, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(ok for _, ok, *_ in res)
    res.sort()
    assert sum(r[2] for r in res) == 4000                  # every vertex owned exactly once
    assert res[0][3] == 0 and res[-1][4] == 0             # chain ends have one neighbour
    assert all(r[4] > 0 for r in res[:-1]) and all(r[3] > 0 for r in res[1:])


def test_slab_partition_is_consistent():
    pg = _global_points(n=3000)
    world = 4
    faces = slab_bounds(0.0, 4.0, world)
    lays = [slab_partition(pg[:, 1], faces, r, width=0.25) for r in range(world)]
    owned = np.concatenate([lay.owned for lay in lays])
    assert np.array_equal(np.sort(owned), np.arange(3000))
    for r in range(world - 1):
        lo, hi = lays[r], lays[r + 1]
        # what r exports upward is exactly what r+1 holds as ghosts from below, same order
        assert np.array_equal(lo.owned[lo.export_above], hi.ghost_below)
        assert np.array_equal(hi.owned[hi.export_below], lo.ghost_above)
        assert np.all(pg[hi.ghost_below, 1] >= faces[r + 1] - 0.25)
        assert np.all(pg[lo.ghost_above, 1] < faces[r + 1] + 0.25)


# ---- variable-size ghost exchange + gather used by generate_mesh_parallel (host message pattern) ----
def _ghost_worker(rank, world, port, q):
    from seismicmesh_b200.parallel import _exchange_ghosts, _gather_points

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dim = 3
        cdev = torch.device("cpu")
        mk = lambda tag, n: np.full((n, dim), 100.0 * rank + tag) + np.arange(n)[:, None]  # noqa: E731
        below = mk(1, 0 if rank == 0 else rank + 1)          # rank r sends r+1 rows down ...
        above = mk(2, 0 if rank == world - 1 else 2 * rank)  # ... and 2r rows up (rank 0: an empty message)
        fa, fb = _exchange_ghosts(below, above, rank, world, dim, cdev)
        ok = True
        if rank < world - 1:  # from above: rank+1's "below" list
            ok &= fa.shape == (rank + 2, dim) and np.array_equal(fa, np.full((rank + 2, dim), 100.0 * (rank + 1) + 1) + np.arange(rank + 2)[:, None])
        else:
            ok &= len(fa) == 0
        if rank > 0:  # from below: rank-1's "above" list
            n = 2 * (rank - 1)
            ok &= fb.shape == (n, dim) and np.array_equal(fb, np.full((n, dim), 100.0 * (rank - 1) + 2) + np.arange(n)[:, None])
        else:
            ok &= len(fb) == 0
        g = _gather_points(mk(3, rank + 2), rank, world, dim, cdev)
        if rank == 0:
            exp = np.vstack([np.full((r + 2, dim), 100.0 * r + 3) + np.arange(r + 2)[:, None] for r in range(world)])
            ok &= np.array_equal(g, exp)
        else:
            ok &= g is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ghost_exchange_and_gather_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ghost_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(ok for _, ok in res)


# ---- the reference's migration / decomp names over torch.distributed (seismicmesh_b200.migration) ----
def _migration_worker(rank, world, port, q):
    import seismicmesh_b200 as sm
    from seismicmesh_b200.parallel import TorchComm, _broadcast_points

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = TorchComm()
        dim = 2
        pts = _global_points(n=600, dim=dim, seed=5)[:, ::-1].copy()  # long along x: `axis=1` cuts x
        ok = True
        # restart path: rank 0's points reach everybody unchanged
        got = _broadcast_points(pts if rank == 0 else None, rank, dim, torch.device("cpu"))
        ok &= np.array_equal(got, pts)
        # blocker on rank 0 + localize_points: every rank gets its block and all the extents
        if rank == 0:
            blocks, extents = sm.decomp.blocker(points=pts, rank=rank, num_blocks=world, axis=1)
        else:
            blocks, extents = None, None
        mine, ext = sm.migration.localize_points(blocks, extents, comm, dim)
        b_all, e_all = sm.decomp.blocker(points=pts, rank=0, num_blocks=world, axis=1)
        ok &= np.array_equal(mine, b_all[rank]) and np.allclose(np.asarray(ext), np.asarray(e_all))
        # exchange: the packed export table of enqueue ([NSB, NSA, ..]; rows (x, y, id)), built by hand here
        nsb, nsa = (0 if rank == 0 else 2 + rank), (0 if rank == world - 1 else 1 + rank)
        exports = np.zeros((1 + nsb + nsa, dim + 1))
        exports[0, :2] = nsb, nsa
        exports[1:, :dim] = 10.0 * rank + np.arange(nsb + nsa)[:, None]
        recv = sm.migration.exchange(comm, rank, world, exports, dim=dim)
        n_exp = (0 if rank == world - 1 else 2 + rank + 1) + (0 if rank == 0 else 1 + rank - 1)
        ok &= (recv.size == 0 and n_exp == 0) or recv.shape == (n_exp, dim)
        # aggregate: local meshes -> rank 0, cells renumbered by the running vertex offset
        lp = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]) + [2.0 * rank, 0.0]
        lt = np.array([[0, 1, 2], [1, 3, 2]])
        gp, gt = sm.migration.aggregate(lp, lt, comm, world, rank, dim=dim)
        if rank == 0:
            ok &= gp.shape == (4 * world, dim) and gt.shape == (2 * world, dim + 1) and gt.max() == 4 * world - 1
            ok &= abs(np.abs(sm.geometry.simp_vol(gp, gt)).sum() - world) < 1e-12
        else:
            ok &= gp is True and gt is True
        ok &= sm.migration.localize_sizing_function("fh", 0.1, None, dim, 1, comm) == "fh"
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_migration_names_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_migration_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(ok for _, ok in res)


def test_blocker_matches_reference():
    from oracle import ref_harness

    if not (ref_harness.reference_available() and ref_harness.native_available()):
        pytest.skip("reference tree only exists in the build container")
    import seismicmesh_b200 as sm

    ref = ref_harness.load_reference()
    for dim, axis, nb in ((2, 0, 3), (2, 1, 4), (3, 0, 2), (3, 1, 5)):
        pts = _global_points(n=900, dim=dim, seed=dim + axis)
        b0, e0 = ref.decomp.blocker(points=pts, rank=0, num_blocks=nb, axis=axis)
        b1, e1 = sm.decomp.blocker(points=pts, rank=0, num_blocks=nb, axis=axis)
        assert len(b0) == len(b1)
        for x, y in zip(b0, b1):
            assert np.array_equal(x, y)
        assert np.allclose(np.asarray(e0), np.asarray(e1), rtol=0, atol=0)
