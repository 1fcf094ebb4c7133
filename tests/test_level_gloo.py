"""Multi-rank host logic on CPU: the box decomposition and the direction-by-direction width-4 halo schedule of
hamers_b200.level (what replaces xfer::RefineSchedule::fillData, RungeKuttaLevelIntegrator.cpp:1568/1701, for
GPU-resident boxes), run with torch.distributed gloo, world_size 2 and 4, numpy slicing standing in for the
CUDA pack/unpack kernels.  Every ghost cell -- faces, edges, corners -- must equal the periodic image."""
import os
import socket

import numpy as np
import pytest

from hamers_b200.level import BoxDecomposition, exchange_halos, exchange_halos_oneshot, neighbour_ranks, oneshot_schedule

G = 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dim, N, ncomp, q, oneshot=False, G=G):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        full = rng.standard_normal((ncomp,) + tuple(reversed(N)))
        dec = BoxDecomposition(dim, N, world, rank)
        n = dec.n
        box = tuple(slice(dec.lo[a], dec.lo[a] + n[a]) for a in reversed(range(dim)))
        U = np.full((ncomp,) + tuple(x + 2 * G for x in reversed(n)), np.nan)
        U[(slice(None),) + tuple(slice(G, -G) for _ in range(dim))] = full[(slice(None),) + box]

        def region(lo, hi):
            return (slice(None),) + tuple(slice(lo[a] + G, hi[a] + G) for a in reversed(range(dim)))

        def pack(lo, hi, buf):
            buf.copy_(torch.from_numpy(np.ascontiguousarray(U[region(lo, hi)]).reshape(-1)))

        def unpack(lo, hi, buf):
            U[region(lo, hi)] = buf.numpy().reshape(U[region(lo, hi)].shape)

        def fill_local(mask):
            for a in range(dim):
                if not (mask >> a) & 1:
                    continue
                ax = dim - a   # numpy axis of direction a (component axis first)
                idx_lo = [slice(None)] * (dim + 1)
                idx_hi = [slice(None)] * (dim + 1)
                src_lo = [slice(None)] * (dim + 1)
                src_hi = [slice(None)] * (dim + 1)
                idx_lo[ax], src_lo[ax] = slice(0, G), slice(n[a], n[a] + G)
                idx_hi[ax], src_hi[ax] = slice(n[a] + G, n[a] + 2 * G), slice(G, 2 * G)
                U[tuple(idx_lo)] = U[tuple(src_lo)]
                U[tuple(idx_hi)] = U[tuple(src_hi)]

        bufs = {}

        def new_buffer(key, numel):
            if key not in bufs:
                bufs[key] = torch.empty(numel, dtype=torch.float64)
            return bufs[key]

        if oneshot:
            def pack_many(boxes, offsets, buf):
                for (lo, hi), off in zip(boxes, offsets):
                    v = np.ascontiguousarray(U[region(lo, hi)]).reshape(-1)
                    buf[off:off + v.size].copy_(torch.from_numpy(v))

            def unpack_many(boxes, offsets, buf):
                for (lo, hi), off in zip(boxes, offsets):
                    shp = U[region(lo, hi)].shape
                    U[region(lo, hi)] = buf[off:off + int(np.prod(shp))].numpy().reshape(shp)

            exchange_halos_oneshot(oneshot_schedule(dec, ncomp, G), ncomp, pack_many, unpack_many, fill_local, new_buffer, dist)
        else:
            exchange_halos(dec.halo_schedule(), ncomp, pack, unpack, fill_local, new_buffer, dist)
        # expected: periodic image of the global array
        idx = [np.arange(dec.lo[a] - G, dec.lo[a] + n[a] + G) % N[a] for a in range(dim)]
        expect = full[(slice(None),) + np.ix_(*reversed(idx))]
        ok = np.array_equal(U, expect)
        q.put((rank, ok, int(np.isnan(U).sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("oneshot", [False, True])
@pytest.mark.parametrize("dim,N,world", [(3, (16, 12, 8), 2), (2, (16, 24), 2), (3, (16, 16, 8), 4)])
def test_halo_exchange_gloo(dim, N, world, oneshot):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dim, N, 3, q, oneshot)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, nans in res:
        assert ok, f"rank {rank}: ghost box differs from the periodic image ({nans} cells never filled)"


@pytest.mark.parametrize("dim,N,world", [(3, (8, 12, 16), 2), (2, (24, 12), 4)])
def test_six_wide_halo_exchange_gloo(dim, N, world):
    """The single-phase schedule with the Navier-Stokes halo width (six ghost cells)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dim, N, 2, q, True, 6)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, nans in res:
        assert ok, f"rank {rank}: ghost box differs from the periodic image ({nans} cells never filled)"


def test_oneshot_schedule_matches_between_ranks():
    """What rank A packs for rank B is, box by box, what B expects from A (NCCL matches messages by order only)."""
    for dim, N, world in ((3, (16, 16, 16), 8), (3, (16, 16, 8), 4), (2, (16, 24), 2), (2, (32, 16), 8)):
        sched = [oneshot_schedule(BoxDecomposition(dim, N, world, r), 5) for r in range(world)]
        for a in range(world):
            for t in sched[a][0]:
                back = [u for u in sched[t.peer][1] if u.peer == a]
                assert len(back) == 1 and back[0].numel == t.numel
                assert [tuple(h - l for l, h in zip(lo, hi)) for lo, hi in t.boxes] == \
                       [tuple(h - l for l, h in zip(lo, hi)) for lo, hi in back[0].boxes]


def test_push_neighbour_table_is_consistent():
    """The fused ghost push stores a cell of rank r near its face in direction o into the ghost box of
    neighbour_ranks(r)[o] at coordinate c - o*n: that cell must be exactly what the periodic level holds there, i.e.
    the neighbour's ghost cell c - o*n maps back to r's interior cell c."""
    for dim, N, world in ((3, (16, 16, 16), 8), (3, (16, 16, 8), 4), (3, (16, 8, 8), 2), (2, (16, 24), 2), (2, (32, 16), 8)):
        decs = [BoxDecomposition(dim, N, world, r) for r in range(world)]
        for d in decs:
            nb = neighbour_ranks(d)
            assert len(nb) == 3 ** dim - 1
            for o, peer in nb.items():
                back = neighbour_ranks(decs[peer])[tuple(-x for x in o)]
                assert back == d.rank
                # a cell at the low corner region of r in direction o, in global periodic coordinates
                c = tuple(0 if o[a] < 0 else (d.n[a] - 1 if o[a] > 0 else 1) for a in range(dim))
                glob = tuple((d.lo[a] + c[a]) % N[a] for a in range(dim))
                ghost = tuple(c[a] - o[a] * d.n[a] for a in range(dim))          # coordinate in the peer's frame
                p = decs[peer]
                assert tuple((p.lo[a] + ghost[a]) % N[a] for a in range(dim)) == glob
                assert any(ghost[a] < 0 or ghost[a] >= p.n[a] for a in range(dim))   # it is a ghost cell there


def test_decomposition_covers_level_once():
    for dim, N, world in ((3, (64, 64, 64), 8), (3, (32, 16, 8), 4), (2, (32, 16), 8)):
        seen = np.zeros(tuple(reversed(N)), dtype=int)
        for r in range(world):
            d = BoxDecomposition(dim, N, world, r)
            box = tuple(slice(d.lo[a], d.lo[a] + d.n[a]) for a in reversed(range(dim)))
            seen[box] += 1
            for a in range(dim):
                for side in (0, 1):
                    nb = BoxDecomposition(dim, N, world, d.neighbour(a, side))
                    assert nb.coords[a] == (d.coords[a] + (1 if side else -1)) % d.grid[a]
        assert (seen == 1).all()


@pytest.mark.parametrize("dim,N,world,ghosts", [(2, (16, 12), 2, 6), (2, (16, 16), 4, 4), (3, (12, 14, 16), 2, 6),
                                                (3, (16, 12, 14), 4, 6), (3, (12, 12, 24), 8, 6), (3, (12, 12, 16), 8, 4)])
def test_push_boxes_fill_every_ghost(dim, N, world, ghosts):
    """push_boxes_of (the table of hb2_push_boxes_dev, the Navier-Stokes level's six-wide direct ghost stores): every
    rank stores its slabs at index - shift into the owner's array, then fills the directions it owns alone locally
    (ghost-inclusive in the exchanged ones): all ghosts -- faces, edges, corners -- equal the periodic image."""
    from hamers_b200.level import push_boxes_of

    g = ghosts
    rng = np.random.default_rng(11)
    full = rng.standard_normal((2,) + tuple(reversed(N)))
    decs = [BoxDecomposition(dim, N, world, r) for r in range(world)]
    arrays = []
    for dec in decs:
        U = np.full((2,) + tuple(x + 2 * g for x in reversed(dec.n)), np.nan)
        box = tuple(slice(dec.lo[a], dec.lo[a] + dec.n[a]) for a in reversed(range(dim)))
        U[(slice(None),) + tuple(slice(g, -g) for _ in range(dim))] = full[(slice(None),) + box]
        arrays.append(U)

    def region(lo, hi):
        return (slice(None),) + tuple(slice(lo[a] + g, hi[a] + g) for a in reversed(range(dim)))

    for dec, U in zip(decs, arrays):
        boxes, peers, shifts = push_boxes_of(dec, g)
        sends = oneshot_schedule(dec, 2, g)[0]
        assert sorted(set(peers)) == [t.peer for t in sends] and len(boxes) == sum(len(t.boxes) for t in sends)
        for (lo, hi), peer, sh in zip(boxes, peers, shifts):
            assert peer != dec.rank
            dlo = tuple(lo[a] - sh[a] for a in range(dim))
            dhi = tuple(hi[a] - sh[a] for a in range(dim))
            assert all(-g <= dlo[a] and dhi[a] <= dec.n[a] + g for a in range(dim))
            arrays[peer][region(dlo, dhi)] = U[region(lo, hi)]
    for dec, U in zip(decs, arrays):
        mask = oneshot_schedule(dec, 2, g)[2]
        for a in range(dim):
            if (mask >> a) & 1:     # ghost-inclusive in the other directions
                ax = U.ndim - 1 - a
                n = dec.n[a]
                idx = (np.arange(-g, n + g) % n) + g
                U[...] = np.take(U, idx, axis=ax)
        want = np.pad(full, [(0, 0)] + [(g, g)] * dim, mode="wrap")
        box = tuple(slice(dec.lo[a], dec.lo[a] + dec.n[a] + 2 * g) for a in reversed(range(dim)))
        assert np.array_equal(U, want[(slice(None),) + box])
