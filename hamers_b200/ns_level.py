"""GPU-resident periodic uniform level of the single-species Navier-Stokes application.

One SSP-RK3 step is three passes of `RungeKuttaLevelIntegrator::advanceLevel`'s stage body
(src/algs/integrator/RungeKuttaLevelIntegrator.cpp:1672-1745) with the patch strategy of
src/apps/Navier-Stokes/NavierStokes.cpp: per stage

    computeFluxesAndSourcesOnPatch   (:1108-1245)  zero source, convective flux + source (WCNS5-JS / HLLC-HLL or the other
                                                   interpolators), diffusive flux ("SIXTH_ORDER")
    advanceSingleStepOnPatch         (:2085-2092)  conservative update with both side fluxes
    same-level ghost fill            (RungeKuttaLevelIntegrator.cpp:1701)

The state carries six ghost cells (max of the convective 4 and the diffusive 6, NavierStokes.cpp ctor) and both
reconstructors read the same arrays (the convective plan is created with num_ghosts = 6), like SAMRAI hands them the
same allocation.  One patch per rank: with torch.distributed initialised the periodic level is split into boxes like the
Euler level (hamers_b200/level.py: BoxDecomposition).  The six-wide halos of a stage's new state are stored straight into
the neighbouring boxes' state arrays (peer-GPU memory opened over CUDA IPC, hb2_push_boxes_dev: one launch, NVLink
stores, no pack / NCCL send-recv / unpack), ordered against the next stage by one tiny stream-ordered all-reduce; the
single-phase NCCL schedule (one message per neighbour, multi-box pack / unpack kernels) fills a freshly set state and is
the whole exchange with push=False (HB2_NS_PUSH=0).  Two routes per stage:

  math = MATH_EXACT  both side fluxes materialised and combined by hb2_advance_stage_ns_dev in the reference's
                     association: a step is bit-identical to the oracle's composition of the same calls;
  math = MATH_FAST   the fused convective stage, then the diffusive divergence accumulated on top
                     (hb2_diffusive_divergence_accumulate_dev): no side flux or source array of either kind is ever
                     written; same terms, re-associated, <= 1e-12.

(tests/test_zz_gpu_diffusive.py; CPU emulation of both routes in tests/test_host_emu_diffusive.py.)"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

from . import abi
from .level import BoxDecomposition, exchange_halos_oneshot, oneshot_schedule, push_boxes_of


class NavierStokesLevel:
    def __init__(self, dim: int, N: Sequence[int], species_gamma: float = 1.4, species_R: float = 1.0,
                 species_mu: float = 1.0e-3, species_mu_v: float = 0.0, species_c_p: float = 3.5, species_Pr: float = 0.72,
                 domain: Tuple[float, float] = (0.0, 1.0), math: int = abi.MATH_EXACT, scheme: int = 0, grid=None,
                 distributed: bool = True, device: str = "cuda", push=None):
        import os

        import torch
        import torch.distributed as dist

        self.torch = torch
        self.dist = dist if (distributed and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None
        nranks = self.dist.get_world_size() if self.dist else 1
        rank = self.dist.get_rank() if self.dist else 0
        self.decomp = BoxDecomposition(dim, tuple(int(x) for x in N[:dim]), nranks, rank, grid)
        self.dim, self.n = dim, self.decomp.n                  # this rank's box
        self.neq = dim + 2
        L = domain[1] - domain[0]
        self.domain_lo = domain[0]
        self.dx = tuple(L / n for n in self.decomp.N)
        c_v = 1.0 / (species_gamma - 1.0) * species_R           # EquationOfStateMixingRulesIdealGas.cpp:119
        self.c_p_eos = species_gamma / (species_gamma - 1.0) * species_R        # :113, the c_p of MAX_DIFFUSIVITY
        self.math = math
        self.cplan = abi.Plan(dim, self.n, species_gamma=(species_gamma,), dx=self.dx, math=math, scheme=scheme,
                              num_ghosts=abi.DIFF_GHOSTS).use_torch_stream()
        self.dplan = abi.DiffusivePlan(dim, self.n, self.dx, species_gamma, c_v, species_mu, species_mu_v, species_c_p,
                                       species_Pr).use_torch_stream()
        if math == abi.MATH_FAST:
            self.dplan.set_math(abi.MATH_FAST)      # the flux-free update in re-associated arithmetic (<= 1e-12)
        self.device = device            # "cuda"; the CPU test suite drives the same loop over emulation-backed plans
        f64 = dict(dtype=torch.float64, device=device)
        g6 = tuple(x + 12 for x in reversed(self.n))
        if push is None:
            push = os.environ.get("HB2_NS_PUSH", "1") != "0"
        self.push = bool(push) and self.dist is not None and device == "cuda"
        self._arrays, self._opened, self.push_tables = [], [], None
        if self.push:
            # IPC-shareable state buffers (U0 and the two intermediate states); the neighbours' are opened below
            self._arrays = [abi.DeviceArray((self.neq,) + g6) for _ in range(3)]
            self.S = [torch.as_tensor(a, device="cuda") for a in self._arrays]
            for t in self.S:
                t.zero_()
        else:
            self.S = [torch.zeros((self.neq,) + g6, **f64) for _ in range(3)]       # U0 and the two intermediate states
        if math == abi.MATH_EXACT:
            self.Fd = [torch.zeros((self.neq,) + self.dplan.side_shape(a), **f64) for a in range(dim)]
            self.Fc = [torch.zeros((self.neq,) + self.dplan.side_shape(a), **f64) for a in range(dim)]
            self.src = torch.zeros((self.neq,) + tuple(reversed(self.n)), **f64)
        self.cur = 0
        self.ghosts_valid = False
        self.alpha, self.beta = abi.SSPRK3_ALPHA, abi.SSPRK3_BETA
        self.oneshot = oneshot_schedule(self.decomp, self.neq, abi.DIFF_GHOSTS)
        self._tables, self._bufs = {}, {}
        if self.dist is not None:
            assert all(n >= abi.DIFF_GHOSTS for n in self.n), "boxes must be at least one halo width wide"
        if self.push:
            self._open_peers(rank, nranks)
        if self.dist is not None:
            self.dist.barrier()     # every rank has opened its neighbours' buffers before anyone steps

    def _open_peers(self, rank, nranks):
        """Per state buffer, the table of hb2_push_boxes_dev: my interior slab next to face / edge / corner o goes to
        the ghost box of the neighbour at o, which sees it shifted by o * n (all boxes of the decomposition are equal)."""
        handles = [None] * nranks
        self.dist.all_gather_object(handles, [abi.ipc_export(a.ptr) for a in self._arrays])
        bases = {rank: [a.ptr for a in self._arrays]}
        boxes, peers, shifts = push_boxes_of(self.decomp, abi.DIFF_GHOSTS)
        for peer in peers:
            if peer not in bases:
                ptrs = [abi.ipc_open(h) for h in handles[peer]]
                self._opened += ptrs
                bases[peer] = ptrs
        self.push_tables = [self.cplan.peer_box_table(boxes, [bases[peer][b] for peer in peers], shifts) for b in range(3)]
        self._flag = self.torch.zeros(1, dtype=self.torch.float32, device="cuda")

    def close(self):
        self.cplan.close()
        self.dplan.close()
        if self._opened or self._arrays:
            self.torch.cuda.synchronize()
            self.dist.barrier()     # nobody unmaps or frees a buffer a neighbour may still be writing into
        for p in self._opened:
            abi.ipc_close(p)
        self._opened = []
        if self._arrays:
            self.S = []
        for a in self._arrays:
            a.free()
        self._arrays = []

    def _interior_slices(self):
        return (slice(None),) + (slice(6, -6),) * self.dim

    def interior(self):
        """View of the interior cells of the current state (writing through it invalidates the ghosts)."""
        self.ghosts_valid = False
        return self.S[self.cur][self._interior_slices()]

    def coordinates(self):
        """cell-centre coordinates of this rank's box per axis (numpy)."""
        return [self.domain_lo + (lo + np.arange(n) + 0.5) * d for lo, n, d in zip(self.decomp.lo, self.n, self.dx)]

    # -- ghost fill: same-level, periodic; across ranks the single-phase schedule of the Euler level, six cells wide ------
    def _buffer(self, key, numel):
        b = self._bufs.get(key)
        if b is None or b.numel() != numel:
            b = self.torch.empty(numel, dtype=self.torch.float64, device=self.device)
            self._bufs[key] = b
        return b

    def _table(self, key, boxes, offsets):
        t = self._tables.get(key)
        if t is None:
            t = self.cplan.box_table(boxes, offsets)
            self._tables[key] = t
        return t

    def fill_ghosts(self, U, buffer=None):
        """Same-level ghost fill of the state U.  buffer = index of U in self.S on the stage path: with push the halos
        are stored straight into the neighbours' copies of that buffer (every rank calls this in the same stage)."""
        if self.dist is None:
            self.dplan.fill_ghosts_periodic(U)
            return
        if self.push and buffer is not None:
            self.cplan.push_boxes(U, self.push_tables[buffer])
            # orders "every rank's stores into my ghosts are complete" before anything reads them (stream-ordered); a
            # rank can be at most one stage ahead of a neighbour, and consecutive stages write different buffers
            self.dist.all_reduce(self._flag)
            if self.oneshot[2]:
                self.dplan.fill_ghosts_periodic(U, self.oneshot[2])
            return
        exchange_halos_oneshot(self.oneshot, self.neq,
                               pack_many=lambda boxes, off, b: self.cplan.pack_boxes(U, self._table("s", boxes, off), b),
                               unpack_many=lambda boxes, off, b: self.cplan.unpack_boxes(U, self._table("r", boxes, off), b),
                               fill_local=lambda mask: self.dplan.fill_ghosts_periodic(U, mask),
                               new_buffer=self._buffer, dist=self.dist)

    @property
    def launch_count(self) -> int:
        return self.cplan.launch_count + self.dplan.launch_count

    def _stage(self, alpha, beta, states, dt, out):
        S = self.S
        newest = S[states[-1]]
        # NavierStokes::computeFluxesAndSourcesOnPatch on the newest state (the RK table only uses its flux)
        if self.math == abi.MATH_EXACT:
            self.dplan.compute_diffusive_flux(newest, dt, self.Fd)
            self.src.zero_()
            self.cplan.compute_flux_and_source(newest, dt, self.Fc, self.src)
            m = len(states)
            none = [None] * (m - 1)
            self.dplan.advance_stage_ns(6, alpha, beta, [S[i] for i in states], none + [self.Fc], none + [self.Fd],
                                        none + [self.src], S[out])
        else:
            self.cplan.fused_stage(alpha, beta, [S[i] for i in states], dt, S[out])
            self.dplan.divergence_accumulate(newest, dt, 6, float(beta[-1]), S[out])
        self.fill_ghosts(S[out], out)

    def stable_dt(self, cfl: float = 1.0) -> float:
        """Level-wide stable time step: NavierStokes::computeSpectralRadiusesAndStableDtOnPatch per box (acoustic radii of
        the convective plan, diffusive radius of the diffusive plan; NavierStokes.cpp:1006-1091), then the MAX all-reduce of
        RungeKuttaLevelIntegrator::getLevelDt over ranks."""
        torch = self.torch
        if not hasattr(self, "_sr"):
            self._sr = torch.zeros(5, dtype=torch.float64, device=self.device)
        if not self.ghosts_valid:
            # the diffusive radius is reduced over the whole six-ghost box like the reference's loop: stale or zero
            # ghosts (rho = 0 -> mu/rho = inf) must not enter it
            self.fill_ghosts(self.S[self.cur])
            self.ghosts_valid = True
        self.cplan.max_wave_speed(self.S[self.cur], self._sr[:4])
        self.dplan.max_spectral_radius(self.S[self.cur], self.c_p_eos, self._sr[4:])
        if self.dist is not None:
            self.dist.all_reduce(self._sr, op=self.dist.ReduceOp.MAX)
        sr = self._sr.cpu().numpy().copy()
        self.spectral_radii = sr
        return cfl / (max(float(sr[4]), float(sr[3])) + 1.0e-15)

    def rk_step(self, dt: float):
        """One SSP-RK3 step (RungeKuttaLevelIntegrator.cpp:3894-3929): three stages, the result becomes the current state."""
        a, b = self.alpha, self.beta
        i0 = self.cur
        i1, i2 = (i0 + 1) % 3, (i0 + 2) % 3
        if not self.ghosts_valid:
            self.fill_ghosts(self.S[i0])
        self._stage(a[0][:1], b[0][:1], [i0], dt, i1)
        self._stage(a[1][:2], b[1][:2], [i0, i1], dt, i2)
        self._stage(a[2][:3], b[2][:3], [i0, i1, i2], dt, i1)      # alpha[2][1] == 0: U1 is free to be overwritten
        self.cur = i1
        self.ghosts_valid = True
