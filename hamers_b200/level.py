"""Uniform periodic level, one GPU-resident box per rank: the host-side mirror of the stage loop of
RungeKuttaLevelIntegrator::advanceLevel (src/algs/integrator/RungeKuttaLevelIntegrator.cpp:1672-1745)
with the ghost fill of xfer::RefineSchedule::fillData (:1568, :1701) replaced by a width-4 halo exchange
between GPU-resident boxes.

Partitioning follows the reference: whole boxes are assigned to ranks (SAMRAI load balancer,
src/exec/main_simulation.hpp:365-422); here the uniform level is cut into a regular process grid, one box
per rank (one rank per GPU).  Per RK stage the newest state's ghosts are filled in ONE phase
(`oneshot_schedule`): every rank packs the face / edge / corner regions of all its (up to 26) neighbours
with one kernel launch (hb2_pack_boxes_dev), exchanges one message per peer with torch.distributed P2P
(NCCL over NVLink on the GPU box), unpacks with one launch, and finally fills the directions it owns
alone from its own periodic image.  The older direction-by-direction schedule (`halo_schedule`: x, then y
including x ghosts, then z including x and y ghosts; 3 x (2 packs + NCCL + 2 unpacks) per stage, measured
at 0.9 ms per stage on 8 GPUs against 3.2 ms of compute) is kept for comparison.  The CPU tests run both
schedules under gloo with numpy slicing standing in for the pack kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

G = 4

PROCESS_GRIDS = {
    # 3-D: x is never split.  Ghost slabs next to an x face are 4- or 6-cell rows (32 / 48 bytes) that make poor NVLink
    # stores; y / z faces are whole rows.  Measured at 512^3 on 8 B200s: (1, 2, 4) 48.2e9 cell-updates/s and
    # Navier-Stokes 34.9e9 against 46.8e9 / 33.0e9 with (2, 2, 2) (profiles/r02_bc_grid_ab_8gpu.txt); the periodic x
    # images are the box's own and are filled locally.
    3: {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (1, 2, 4)},
    2: {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)},
}


def _grid_override():
    """HB2_PROCESS_GRID="gx,gy,gz" (or "gx,gy"): replaces the default process grid of that many ranks (measurements of
    other box shapes; every rank of a job must see the same value)."""
    import os

    env = os.environ.get("HB2_PROCESS_GRID", "")
    if env:
        g = tuple(int(x) for x in env.split(","))
        PROCESS_GRIDS[len(g)][int(np.prod(g))] = g


_grid_override()


@dataclass
class Message:
    peer: int
    lo: Tuple[int, ...]
    hi: Tuple[int, ...]
    tag: int


@dataclass
class Phase:
    direction: int
    local_periodic: bool
    sends: List[Message]
    recvs: List[Message]


class BoxDecomposition:
    """Regular block decomposition of a periodic uniform level of N cells over a process grid."""

    def __init__(self, dim: int, N: Sequence[int], nranks: int, rank: int, grid: Optional[Sequence[int]] = None):
        self.dim = dim
        self.N = tuple(int(x) for x in N)
        self.nranks = nranks
        self.rank = rank
        self.grid = tuple(grid) if grid is not None else PROCESS_GRIDS[dim][nranks]
        assert int(np.prod(self.grid)) == nranks
        for a in range(dim):
            assert self.N[a] % self.grid[a] == 0, "level dims must divide over the process grid"
        self.n = tuple(self.N[a] // self.grid[a] for a in range(dim))
        self.coords = self.coords_of(rank)
        self.lo = tuple(self.coords[a] * self.n[a] for a in range(dim))

    def coords_of(self, rank: int):
        c = []
        for a in range(self.dim):
            c.append(rank % self.grid[a])
            rank //= self.grid[a]
        return tuple(c)

    def rank_of(self, coords):
        r, mul = 0, 1
        for a in range(self.dim):
            r += (coords[a] % self.grid[a]) * mul
            mul *= self.grid[a]
        return r

    def neighbour(self, direction: int, side: int) -> int:
        c = list(self.coords)
        c[direction] += -1 if side == 0 else 1
        return self.rank_of(c)

    def halo_schedule(self) -> List[Phase]:
        """Direction-by-direction exchange.  Tangential extent: ghost-inclusive in already-exchanged
        directions, interior in the remaining ones."""
        phases = []
        for d in range(self.dim):
            if self.grid[d] == 1:
                phases.append(Phase(d, True, [], []))
                continue
            tlo = [(-G if a < d else 0) for a in range(self.dim)]
            thi = [(self.n[a] + G if a < d else self.n[a]) for a in range(self.dim)]

            def box(lo_d, hi_d):
                lo, hi = list(tlo), list(thi)
                lo[d], hi[d] = lo_d, hi_d
                return tuple(lo), tuple(hi)

            nlo, nhi = self.neighbour(d, 0), self.neighbour(d, 1)
            n = self.n[d]
            sends = [Message(nlo, *box(0, G), tag=2 * d + 0),          # my low slab -> low neighbour's high ghost
                     Message(nhi, *box(n - G, n), tag=2 * d + 1)]      # my high slab -> high neighbour's low ghost
            recvs = [Message(nhi, *box(n, n + G), tag=2 * d + 0),      # high neighbour's low slab
                     Message(nlo, *box(-G, 0), tag=2 * d + 1)]         # low neighbour's high slab
            phases.append(Phase(d, False, sends, recvs))
        return phases


@dataclass
class PeerTraffic:
    """Everything one rank exchanges with one peer in a ghost fill: boxes in a canonical order (both sides sort by
    the direction code of the SENDER), packed back to back into one message."""
    peer: int
    boxes: List[Tuple[Tuple[int, ...], Tuple[int, ...]]]
    numel: int = 0


def _code(o):
    return sum((o[a] + 1) * 3 ** a for a in range(len(o)))


def neighbour_ranks(dec: "BoxDecomposition"):
    """{offset: rank} of the (up to 26, 8 in 2-D) boxes around this rank's box on the periodic process grid; a rank
    is its own neighbour in the directions it owns alone."""
    import itertools

    return {o: dec.rank_of([dec.coords[a] + o[a] for a in range(dec.dim)])
            for o in itertools.product((-1, 0, 1), repeat=dec.dim) if any(o)}


def oneshot_schedule(dec: "BoxDecomposition", ncomp: int, ghosts: int = G):
    """Single-phase ghost fill: every rank sends each of its (up to 26) neighbours the face / edge / corner region
    that neighbour's ghost box needs, all neighbours at once, ONE message per peer.  Directions owned by a single rank
    are periodic images of the rank's own box and are filled locally afterwards (`local_mask`), ghost-inclusive in the
    exchanged directions.  `ghosts` is the halo width (4 for the convective path, 6 for a Navier-Stokes state).
    Returns (sends, recvs, local_mask) with sends / recvs lists of PeerTraffic."""
    import itertools

    dim, n, grid = dec.dim, dec.n, dec.grid
    local_mask = sum(1 << a for a in range(dim) if grid[a] == 1)
    sends, recvs = {}, {}
    for o in itertools.product((-1, 0, 1), repeat=dim):
        if not any(o) or any(o[a] != 0 and grid[a] == 1 for a in range(dim)):
            continue
        peer = dec.rank_of([dec.coords[a] + o[a] for a in range(dim)])
        # my interior cells next to the face / edge / corner in direction o
        slo = tuple(n[a] - ghosts if o[a] > 0 else 0 for a in range(dim))
        shi = tuple(ghosts if o[a] < 0 else n[a] for a in range(dim))
        sends.setdefault(peer, []).append((_code(o), slo, shi))
        # my ghost cells in direction o, sent by the neighbour there towards -o
        rlo = tuple(-ghosts if o[a] < 0 else (n[a] if o[a] > 0 else 0) for a in range(dim))
        rhi = tuple(0 if o[a] < 0 else (n[a] + ghosts if o[a] > 0 else n[a]) for a in range(dim))
        recvs.setdefault(peer, []).append((_code(tuple(-x for x in o)), rlo, rhi))

    def finish(table):
        out = []
        for peer in sorted(table):
            boxes = [(lo, hi) for _, lo, hi in sorted(table[peer])]
            numel = sum(ncomp * int(np.prod([h - l for l, h in zip(lo, hi)])) for lo, hi in boxes)
            out.append(PeerTraffic(peer, boxes, numel))
        return out

    return finish(sends), finish(recvs), local_mask


def push_boxes_of(dec: "BoxDecomposition", ghosts: int = G):
    """The sends of oneshot_schedule as direct stores (hb2_push_boxes_dev): (boxes, peers, shifts) -- box b = my
    interior cells next to the face / edge / corner in direction o, owned as ghosts by rank peers[b], whose box sees
    them at cell index - shifts[b] (shifts[b] = o * n; every box of the decomposition has the same size)."""
    import itertools

    dim, n, grid = dec.dim, dec.n, dec.grid
    boxes, peers, shifts = [], [], []
    for o in itertools.product((-1, 0, 1), repeat=dim):
        if not any(o) or any(o[a] != 0 and grid[a] == 1 for a in range(dim)):
            continue
        peers.append(dec.rank_of([dec.coords[a] + o[a] for a in range(dim)]))
        boxes.append((tuple(n[a] - ghosts if o[a] > 0 else 0 for a in range(dim)),
                      tuple(ghosts if o[a] < 0 else n[a] for a in range(dim))))
        shifts.append(tuple(o[a] * n[a] for a in range(dim)))
    return boxes, peers, shifts


def exchange_halos_oneshot(schedule, ncomp: int, pack_many: Callable, unpack_many: Callable, fill_local: Callable,
                           new_buffer: Callable, dist):
    """Run the single-phase ghost fill: pack_many(boxes, offsets, buf) / unpack_many(boxes, offsets, buf) move a list of
    boxes to / from positions `offsets` (in elements) of ONE buffer -- one kernel launch each on the GPU."""
    sends, recvs, local_mask = schedule
    if sends:
        def layout(traffic):
            boxes, offsets, spans, pos = [], [], [], 0
            for t in traffic:
                spans.append((t.peer, pos, t.numel))
                for lo, hi in t.boxes:
                    boxes.append((lo, hi))
                    offsets.append(pos)
                    pos += ncomp * int(np.prod([h - l for l, h in zip(lo, hi)]))
            return boxes, offsets, spans, pos

        sboxes, soff, sspans, stotal = layout(sends)
        rboxes, roff, rspans, rtotal = layout(recvs)
        sbuf = new_buffer("send", stotal)
        rbuf = new_buffer("recv", rtotal)
        pack_many(sboxes, soff, sbuf)
        ops = [dist.P2POp(dist.isend, sbuf[p0:p0 + m], peer) for peer, p0, m in sspans]
        ops += [dist.P2POp(dist.irecv, rbuf[p0:p0 + m], peer) for peer, p0, m in rspans]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        unpack_many(rboxes, roff, rbuf)
    if local_mask:
        fill_local(local_mask)


def exchange_halos(phases: List[Phase], ncomp: int,
                   pack: Callable, unpack: Callable, fill_local: Callable,
                   new_buffer: Callable, dist, device_sync: Optional[Callable] = None):
    """Run the halo schedule.  pack(lo, hi, buf), unpack(lo, hi, buf), fill_local(mask) act on the state;
    new_buffer(numel) allocates a message buffer; dist is torch.distributed (or a stand-in)."""
    local_mask = 0
    for ph in phases:
        if ph.local_periodic:
            local_mask |= 1 << ph.direction
    for ph in phases:
        if ph.local_periodic:
            fill_local(1 << ph.direction)
            continue
        ops, rbufs, keep = [], [], []
        for m in ph.sends:
            numel = ncomp * int(np.prod([h - l for l, h in zip(m.lo, m.hi)]))
            b = new_buffer(("s", ph.direction, m.tag), numel)
            pack(m.lo, m.hi, b)
            keep.append(b)
            ops.append(dist.P2POp(dist.isend, b, m.peer, tag=m.tag))
        for m in ph.recvs:
            numel = ncomp * int(np.prod([h - l for l, h in zip(m.lo, m.hi)]))
            b = new_buffer(("r", ph.direction, m.tag), numel)
            rbufs.append((m, b))
            ops.append(dist.P2POp(dist.irecv, b, m.peer, tag=m.tag))
        if device_sync is not None:
            device_sync()
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for m, b in rbufs:
            unpack(m.lo, m.hi, b)


class UniformLevel:
    """GPU-resident periodic uniform level advanced with SSP-RK3 fused stages.

    push=True (default): the ghost fill of every new state is FUSED INTO THE LAST SWEEP of the stage that produces it
    (hb2_fused_stage_push_dev): cells within the ghost width of a box face are stored straight into the ghost boxes
    of the neighbouring boxes -- peer-GPU memory opened over CUDA IPC, the stores travel over NVLink -- and into the
    box's own ghosts where it is its own periodic neighbour.  No pack / NCCL send-recv / unpack and no separate
    periodic-fill kernel remain on the stage path; ranks only meet in one tiny all-reduce per stage, which orders
    "all pushes of stage s have landed" before "stage s+1 reads its ghosts".  push=False keeps the explicit
    exchange (one-shot NCCL schedule), which is also what fills the ghosts of a freshly set state."""

    def __init__(self, dim: int, N: Sequence[int], flow_model: int = 0, species_gamma: Sequence[float] = (1.4,),
                 domain: Tuple[float, float] = (-1.0, 1.0), math: int = 1, weno_p: int = 2,
                 grid: Optional[Sequence[int]] = None, push: Optional[bool] = None, scheme: int = 0,
                 distributed: bool = True, species_R: Sequence[float] = ()):
        import itertools
        import os

        import torch
        import torch.distributed as dist

        from . import abi

        self.torch = torch
        # distributed=False: the whole level as one box on this GPU even inside a torchrun job (the N = 1 replica that
        # multi-GPU runs are compared with)
        self.dist = dist if (distributed and dist.is_available() and dist.is_initialized()) else None
        nranks = self.dist.get_world_size() if self.dist else 1
        rank = self.dist.get_rank() if self.dist else 0
        self.decomp = BoxDecomposition(dim, N, nranks, rank, grid)
        self.dim = dim
        self.dx = tuple((domain[1] - domain[0]) / n for n in self.decomp.N)
        self.plan = abi.Plan(dim, self.decomp.n, flow_model=flow_model, species_gamma=species_gamma, dx=self.dx,
                             weno_p=weno_p, math=math, scheme=scheme, species_R=species_R).use_torch_stream()
        self.ncomp, self.neq = self.plan.ncomp, self.plan.neq
        shape = (self.ncomp,) + self.plan.ghost_shape
        if push is None:
            # measured on B200: on ONE box the separate periodic-fill kernel (0.16 ms at 512^3) is cheaper than the
            # pushes (+0.35 ms on the last sweep); across boxes the push replaces pack + NCCL + unpack (~0.5 ms)
            env = os.environ.get("HB2_LEVEL_PUSH", "")
            push = (nranks > 1) if env == "" else (env != "0")
        # the push needs every box to be at least one ghost width wide
        self.push = bool(push) and all(n >= G for n in self.decomp.n)
        self._arrays, self._opened = [], []
        if self.push and self.dist is not None:
            # IPC-shareable state buffers; every rank opens the buffers of its (up to 26) neighbours
            self._arrays = [abi.DeviceArray(shape) for _ in range(3)]
            self.S = [torch.as_tensor(a, device="cuda") for a in self._arrays]
            for t in self.S:
                t.zero_()
            mine = [abi.ipc_export(a.ptr) for a in self._arrays]
            handles = [None] * nranks
            self.dist.all_gather_object(handles, mine)
            bases = {rank: [a.ptr for a in self._arrays]}
        else:
            self.S = [torch.zeros(shape, dtype=torch.float64, device="cuda") for _ in range(3)]
            bases = {rank: [t.data_ptr() for t in self.S]}
        self.push_tables = None
        if self.push:
            per_buffer = [dict() for _ in range(3)]
            for o, peer in neighbour_ranks(self.decomp).items():
                if peer not in bases:
                    ptrs = [abi.ipc_open(h) for h in handles[peer]]
                    self._opened += ptrs
                    bases[peer] = ptrs
                for b in range(3):
                    per_buffer[b][o] = bases[peer][b]
            self.push_tables = [self.plan.push_table(t) for t in per_buffer]
            self._flag = torch.zeros(1, dtype=torch.float32, device="cuda")
        self._force_fill = os.environ.get("HB2_LEVEL_FORCE_FILL", "0") == "1"
        self.ghosts_valid = False
        self.cur = 0
        self.phases = self.decomp.halo_schedule()
        self.oneshot = oneshot_schedule(self.decomp, self.ncomp)
        self._tables = {}
        self._bufs = {}
        self.alpha = abi.SSPRK3_ALPHA
        self.beta = abi.SSPRK3_BETA
        self.time = 0.0
        if self.dist is not None:
            self.dist.barrier()   # every rank has opened its neighbours' buffers before anyone steps

    # -- state access ---------------------------------------------------------------------
    def _interior_slices(self):
        return (slice(None),) + tuple(slice(G, -G) for _ in range(self.dim))

    def set_interior(self, U):
        """U: (ncomp, *cell_shape) numpy array or CUDA tensor of THIS rank's box."""
        t = self.torch.as_tensor(U, dtype=self.torch.float64).to("cuda")
        self.S[self.cur][self._interior_slices()] = t
        self.ghosts_valid = False

    def interior(self):
        """View of this rank's interior cells (writing through it invalidates the ghosts: they are refilled by the
        explicit exchange before the next step)."""
        self.ghosts_valid = False
        return self.S[self.cur][self._interior_slices()]

    def local_coordinates(self, domain_lo=-1.0):
        """cell-centre coordinates of this rank's box per axis (numpy)."""
        return [domain_lo + (self.decomp.lo[a] + np.arange(self.decomp.n[a]) + 0.5) * self.dx[a] for a in range(self.dim)]

    # -- ghost fill --------------------------------------------------------------------------
    def _buffer(self, key, numel):
        b = self._bufs.get(key)
        if b is None or b.numel() != numel:
            b = self.torch.empty(numel, dtype=self.torch.float64, device="cuda")
            self._bufs[key] = b
        return b

    def fill_ghosts(self, U):
        if self.dist is None:
            self.plan.fill_ghosts_periodic(U, (1 << self.dim) - 1)
            return
        exchange_halos_oneshot(self.oneshot, self.ncomp,
                               pack_many=lambda boxes, off, b: self.plan.pack_boxes(U, self._table("s", boxes, off), b),
                               unpack_many=lambda boxes, off, b: self.plan.unpack_boxes(U, self._table("r", boxes, off), b),
                               fill_local=lambda mask: self.plan.fill_ghosts_periodic(U, mask),
                               new_buffer=self._buffer, dist=self.dist)

    def _table(self, key, boxes, offsets):
        t = self._tables.get(key)
        if t is None:
            t = self.plan.box_table(boxes, offsets)
            self._tables[key] = t
        return t

    # -- time stepping -----------------------------------------------------------------------
    def _stage(self, alpha, beta, states, dt, out):
        """One fused stage writing S[out]; with push the ghosts of S[out] are valid on every rank afterwards."""
        S = self.S
        if self.push:
            self.plan.fused_stage(alpha, beta, [S[i] for i in states], dt, S[out], push=self.push_tables[out])
            if self.dist is not None:
                # orders "every rank's pushes into my ghosts are complete" before the next stage (stream-ordered)
                self.dist.all_reduce(self._flag)
            if self._force_fill:      # diagnostic (HB2_LEVEL_FORCE_FILL=1): time the push with the ghosts kept valid regardless
                self.fill_ghosts(S[out])
        else:
            self.plan.fused_stage(alpha, beta, [S[i] for i in states], dt, S[out])
            self.fill_ghosts(S[out])

    def rk_step(self, dt: float):
        """One SSP-RK3 step = three passes of the hot path (fused flux / update / ghost fill of the new state)."""
        a, b = self.alpha, self.beta
        i0 = self.cur
        i1, i2 = (i0 + 1) % 3, (i0 + 2) % 3
        if not self.ghosts_valid:
            self.fill_ghosts(self.S[i0])
        # stage 0: U1 = U0 + L(U0)
        self._stage(a[0][:1], b[0][:1], [i0], dt, i1)
        # stage 1: U2 = 3/4 U0 + 1/4 U1 + 1/4 L(U1)
        self._stage(a[1][:2], b[1][:2], [i0, i1], dt, i2)
        # stage 2: U3 = 1/3 U0 + 2/3 U2 + 2/3 L(U2), written over U1 (alpha[2][1] == 0)
        self._stage(a[2][:3], b[2][:3], [i0, i1, i2], dt, i1)
        self.cur = i1
        self.ghosts_valid = True
        self.time += dt

    def rk_step_host(self, host, dt: float):
        """One SSP-RK3 step on HOST memory: `host` is this rank's box, interior cells only, (ncomp, *cell_shape), pinned
        torch tensor.  H2D of the box, same-level ghost fill across ranks, three fused stages, D2H of the new interior
        (the per-rank form of hb2_advance_level_host when the level is spread over several GPUs)."""
        torch = self.torch
        self.S[self.cur][self._interior_slices()].copy_(host, non_blocking=True)
        self.ghosts_valid = False
        self.rk_step(dt)
        host.copy_(self.S[self.cur][self._interior_slices()], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def stable_dt(self, cfl: float = 1.0) -> float:
        """Level-wide stable time step: Euler::computeSpectralRadiusesAndStableDtOnPatch per box, then the MAX
        all-reduce of RungeKuttaLevelIntegrator::getLevelDt (RungeKuttaLevelIntegrator.cpp:1786-1928) over ranks."""
        torch = self.torch
        if not hasattr(self, "_sr"):
            self._sr = torch.zeros(4, dtype=torch.float64, device="cuda")
        self.plan.max_wave_speed(self.S[self.cur], self._sr)
        if self.dist is not None:
            self.dist.all_reduce(self._sr, op=self.dist.ReduceOp.MAX)
        self.spectral_radii = self._sr.cpu().numpy().copy()
        # Euler.cpp:846-861: dt = 1 / (spectral radius + HAMERS_EPSILON)
        return cfl / (float(self.spectral_radii[3]) + 1.0e-15)

    def advance(self, dt: float, nsteps: int):
        for _ in range(nsteps):
            self.rk_step(dt)

    def close(self):
        from . import abi

        self.plan.close()
        if self.dist is not None and (self._opened or self._arrays):
            self.torch.cuda.synchronize()
            self.dist.barrier()   # nobody unmaps or frees a buffer a neighbour may still be writing into
        for p in self._opened:
            abi.ipc_close(p)
        self._opened = []
        self.S = []
        for a in self._arrays:
            a.free()
        self._arrays = []
