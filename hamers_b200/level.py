"""Uniform periodic level, one GPU-resident box per rank: the host-side mirror of the stage loop of
RungeKuttaLevelIntegrator::advanceLevel (src/algs/integrator/RungeKuttaLevelIntegrator.cpp:1672-1745)
with the ghost fill of xfer::RefineSchedule::fillData (:1568, :1701) replaced by a width-4 halo exchange
between GPU-resident boxes.

Partitioning follows the reference: whole boxes are assigned to ranks (SAMRAI load balancer,
src/exec/main_simulation.hpp:365-422); here the uniform level is cut into a regular process grid, one box
per rank (one rank per GPU).  Per RK stage the newest state's ghosts are filled direction by direction
(x, then y including x ghosts, then z including x and y ghosts) so that the edge cells the shock sensor
needs are valid; a direction with a single rank is a local periodic copy.  Messages are packed / unpacked
by the CUDA kernels behind hb2_pack_box_dev / hb2_unpack_box_dev and moved with torch.distributed P2P
(NCCL over NVLink on the GPU box; gloo in the CPU tests, which substitute numpy slicing for the pack
kernels to exercise exactly this schedule).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

G = 4

PROCESS_GRIDS = {
    3: {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)},
    2: {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)},
}


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
    """GPU-resident periodic uniform level advanced with SSP-RK3 fused stages."""

    def __init__(self, dim: int, N: Sequence[int], flow_model: int = 0, species_gamma: Sequence[float] = (1.4,),
                 domain: Tuple[float, float] = (-1.0, 1.0), math: int = 1, weno_p: int = 2,
                 grid: Optional[Sequence[int]] = None):
        import torch
        import torch.distributed as dist

        from . import abi

        self.torch = torch
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        nranks = self.dist.get_world_size() if self.dist else 1
        rank = self.dist.get_rank() if self.dist else 0
        self.decomp = BoxDecomposition(dim, N, nranks, rank, grid)
        self.dim = dim
        self.dx = tuple((domain[1] - domain[0]) / n for n in self.decomp.N)
        self.plan = abi.Plan(dim, self.decomp.n, flow_model=flow_model, species_gamma=species_gamma, dx=self.dx,
                             weno_p=weno_p, math=math).use_torch_stream()
        self.ncomp, self.neq = self.plan.ncomp, self.plan.neq
        shape = (self.ncomp,) + self.plan.ghost_shape
        self.S = [torch.zeros(shape, dtype=torch.float64, device="cuda") for _ in range(3)]
        self.cur = 0
        self.phases = self.decomp.halo_schedule()
        self._bufs = {}
        self.alpha = abi.SSPRK3_ALPHA
        self.beta = abi.SSPRK3_BETA
        self.time = 0.0

    # -- state access ---------------------------------------------------------------------
    def _interior_slices(self):
        return (slice(None),) + tuple(slice(G, -G) for _ in range(self.dim))

    def set_interior(self, U):
        """U: (ncomp, *cell_shape) numpy array or CUDA tensor of THIS rank's box."""
        t = self.torch.as_tensor(U, dtype=self.torch.float64).to("cuda")
        self.S[self.cur][self._interior_slices()] = t

    def interior(self):
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
        exchange_halos(self.phases, self.ncomp,
                       pack=lambda lo, hi, b: self.plan.pack_box(U, lo, hi, b),
                       unpack=lambda lo, hi, b: self.plan.unpack_box(U, lo, hi, b),
                       fill_local=lambda mask: self.plan.fill_ghosts_periodic(U, mask),
                       new_buffer=self._buffer, dist=self.dist)

    # -- time stepping -----------------------------------------------------------------------
    def rk_step(self, dt: float):
        """One SSP-RK3 step = three passes of the hot path (ghost fill + fused flux/update)."""
        a, b = self.alpha, self.beta
        i0 = self.cur
        i1, i2 = (i0 + 1) % 3, (i0 + 2) % 3
        S = self.S
        # stage 0: U1 = U0 + L(U0)
        self.fill_ghosts(S[i0])
        self.plan.fused_stage(a[0][:1], b[0][:1], [S[i0]], dt, S[i1])
        # stage 1: U2 = 3/4 U0 + 1/4 U1 + 1/4 L(U1)
        self.fill_ghosts(S[i1])
        self.plan.fused_stage(a[1][:2], b[1][:2], [S[i0], S[i1]], dt, S[i2])
        # stage 2: U3 = 1/3 U0 + 2/3 U2 + 2/3 L(U2), written over U1 (alpha[2][1] == 0)
        self.fill_ghosts(S[i2])
        self.plan.fused_stage(a[2][:3], b[2][:3], [S[i0], S[i1], S[i2]], dt, S[i1])
        self.cur = i1
        self.time += dt

    def advance(self, dt: float, nsteps: int):
        for _ in range(nsteps):
            self.rk_step(dt)

    def close(self):
        self.plan.close()
