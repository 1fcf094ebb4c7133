"""GPU-resident two-level patch hierarchy around the convective hot path (BASELINE.json config 4: "2D
multi-species Richtmyer-Meshkov shock-interface with 2-level patch AMR").

Host-side mirror, for ONE coarse patch covering the domain and ONE fine patch refined by `ratio` over the coarse cells
[clo, chi), of what the reference does around the per-patch path when a finer level exists (path:line under the reference tree):

  RungeKuttaLevelIntegrator::advanceLevel              src/algs/integrator/RungeKuttaLevelIntegrator.cpp:1457-1929
      ghost fill at the current time (same level, coarser level with time interpolation, physical boundary), zeroed flux /
      source sums, per stage: same-level fill that leaves the coarse-fine ghost values of the first fill in place (:1672-1745),
      computeFluxesAndSourcesOnPatch + advanceSingleStepOnPatch with the gamma-weighted flux sums (Euler.cpp:1555-1640)
  postprocessFluxAndSourceData                         :2968-3230 (algs_upfluxsum{2,3}d.f)  -> hb2_amr_fluxsum_update_dev
  synchronizeLevelWithCoarser                          :2131-2209
      coarsen flux integrals onto the coarse flux      -> hb2_amr_coarsen_fluxsum_dev
      Euler::synchronizeFluxes (Euler.cpp:1682-1949)   -> hb2_advance_stage_dev(alpha = beta = 1) on the accumulated flux
      conservative coarsen of the solution             -> hb2_amr_coarsen_dev
  SAMRAI's TimeRefinementIntegrator order: the coarse level steps first, then `ratio` fine steps, then the synchronisation.

Both levels keep their state, the three stage states, the stage flux and the flux / source sums in HBM; every operator is a
kernel of the C-ABI library (include/hamers_b200.h); torch only owns the memory and a few strided ghost-slab copies.  The
levels run the materialised-flux route of the hot path (the flux integrals need the side fluxes, like the reference).
Regridding / tagging / load balancing are SAMRAI's and out of scope: the fine box is static."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import abi

G = abi.GHOSTS


class _Level:
    """State and work arrays of one level (one patch)."""

    def __init__(self, torch, plan, device):
        self.plan = plan
        f64 = dict(dtype=torch.float64, device=device)
        dim, neq, ncomp = plan.dim, plan.neq, plan.ncomp
        self.S = [torch.zeros((ncomp,) + plan.ghost_shape, **f64) for _ in range(4)]        # stage states U^(0..3)
        self.F = [torch.zeros((neq,) + plan.side_shape(d), **f64) for d in range(dim)]       # flux of the current stage
        self.Facc = [torch.zeros((neq,) + plan.side_shape(d), **f64) for d in range(dim)]    # gamma-weighted sums of a step
        self.src = torch.zeros((neq,) + plan.cell_shape, **f64)
        self.Sacc = torch.zeros((neq,) + plan.cell_shape, **f64)
        self.cur = 0


class TwoLevelHierarchy:
    def __init__(self, dim: int, N: Sequence[int], clo: Sequence[int], chi: Sequence[int], ratio: int = 2,
                 periodic: Sequence[bool] = (True, True, True), flow_model: int = abi.SINGLE_SPECIES,
                 species_gamma: Sequence[float] = (1.4,), species_R: Sequence[float] = (), dx: Sequence[float] = (1.0, 1.0, 1.0),
                 math: int = abi.MATH_EXACT, scheme: int = 0, device: str = "cuda", plan_factory=None, pair_factory=None):
        import torch

        self.torch, self.dim, self.device = torch, dim, device
        self.N = tuple(int(x) for x in N[:dim])
        self.clo, self.chi = tuple(int(x) for x in clo[:dim]), tuple(int(x) for x in chi[:dim])
        self.r = int(ratio)
        self.periodic = tuple(bool(p) for p in periodic[:dim])
        self.dxc = tuple(float(x) for x in dx[:dim])
        self.dxf = tuple(h / self.r for h in self.dxc)
        self.nf = tuple(self.r * (self.chi[a] - self.clo[a]) for a in range(dim))
        for a in range(dim):
            assert 0 <= self.clo[a] < self.chi[a] <= self.N[a], "the fine box must lie inside the coarse level"
        plan_factory = plan_factory or abi.Plan
        pair_factory = pair_factory or abi.AmrPair
        kw = dict(flow_model=flow_model, species_gamma=species_gamma, species_R=species_R, math=math, scheme=scheme)
        self.coarse = _Level(torch, plan_factory(dim, self.N, dx=self.dxc, **kw).use_torch_stream(), device)
        self.fine = _Level(torch, plan_factory(dim, self.nf, dx=self.dxf, **kw).use_torch_stream(), device)
        pc = self.coarse.plan
        self.ncomp, self.neq = pc.ncomp, pc.neq
        self.pair = pair_factory(dim, self.N, self.nf, (self.r,) * dim, self.clo, self.dxc, self.dxf, self.ncomp, self.neq)
        # a direction the fine patch spans completely AND that is periodic is filled from the fine patch itself; a fine face
        # on a non-periodic domain boundary is a physical boundary of the fine level too
        self.self_periodic = tuple(self.periodic[a] and self.clo[a] == 0 and self.chi[a] == self.N[a] for a in range(dim))
        self.fine_physical = tuple((not self.periodic[a] and self.clo[a] == 0, not self.periodic[a] and self.chi[a] == self.N[a])
                                   for a in range(dim))
        self.cf_boxes = self._cf_boxes()
        f64 = dict(dtype=torch.float64, device=device)
        self.fluxsum = [torch.zeros((self.neq,) + tuple(self.nf[a] for a in reversed(range(dim)) if a != d), **f64)
                        for d in range(dim) for _ in (0, 1)]
        self.Uc_old = torch.zeros_like(self.coarse.S[0])
        self.alpha, self.beta, self.gamma = abi.SSPRK3_ALPHA, abi.SSPRK3_BETA, abi.SSPRK3_GAMMA
        self.time = 0.0

    def close(self):
        self.coarse.plan.close()
        self.fine.plan.close()

    # -- state access ---------------------------------------------------------------------------------------------------------
    def _inner(self):
        return (slice(None),) + (slice(G, -G),) * self.dim

    def _box(self, lo, hi):
        return (slice(None),) + tuple(slice(lo[a] + G, hi[a] + G) for a in reversed(range(self.dim)))

    @property
    def Uc(self):
        return self.coarse.S[self.coarse.cur]

    @property
    def Uf(self):
        return self.fine.S[self.fine.cur]

    def set_coarse(self, U):
        self.Uc[self._inner()] = self.torch.as_tensor(U, dtype=self.torch.float64).to(self.device)

    def set_fine(self, U):
        self.Uf[self._inner()] = self.torch.as_tensor(U, dtype=self.torch.float64).to(self.device)

    def coarsen_fine_onto_coarse(self):
        self.pair.coarsen(self.Uf, self.clo, self.chi, self.Uc)

    def initialize_fine_from_coarse(self):
        self._fill_coarse(self.Uc)
        self.pair.refine(self.Uc, None, 0.0, (0,) * self.dim, self.nf, self.Uf)

    # -- ghost fills ------------------------------------------------------------------------------------------------------------
    def _fill_coarse(self, U):
        p = self.coarse.plan
        mask = 0
        for a in range(self.dim):
            if self.periodic[a]:
                mask |= 1 << a
            else:
                p.fill_ghosts_extrapolate(U, a, 0)
                p.fill_ghosts_extrapolate(U, a, 1)
        if mask:
            p.fill_ghosts_periodic(U, mask)

    def _cf_boxes(self):
        dim, nf = self.dim, self.nf
        boxes = []
        for a in range(dim):
            if self.self_periodic[a]:
                continue
            for side in (0, 1):
                if self.fine_physical[a][side]:
                    continue
                lo, hi = [], []
                for b in range(dim):
                    if b == a:
                        lo.append(-G if side == 0 else nf[b])
                        hi.append(0 if side == 0 else nf[b] + G)
                    elif self.self_periodic[b]:
                        lo.append(0)
                        hi.append(nf[b])
                    else:
                        lo.append(0 if self.fine_physical[b][0] else -G)
                        hi.append(nf[b] if self.fine_physical[b][1] else nf[b] + G)
                boxes.append((tuple(lo), tuple(hi)))
        return boxes

    def _fill_fine_same_level(self, U):
        p = self.fine.plan
        mask = 0
        for a in range(self.dim):
            for side in (0, 1):
                if self.fine_physical[a][side]:
                    p.fill_ghosts_extrapolate(U, a, side)
            if self.self_periodic[a]:
                mask |= 1 << a
        if mask:
            p.fill_ghosts_periodic(U, mask)

    def _fill_fine_from_coarse(self, U, Uc_old, Uc_new, tfrac):
        for lo, hi in self.cf_boxes:
            self.pair.refine(Uc_old, Uc_new, tfrac, lo, hi, U)
        self._fill_fine_same_level(U)

    def _coarse_stage_fill(self, U, U0):
        self._fill_coarse(U)

    def _fine_stage_fill(self, U, U0):
        for lo, hi in self.cf_boxes:
            s = self._box(lo, hi)
            U[s] = U0[s]                     # "Dirichlet": the coarse-fine ghost values of the first fill stay in place
        self._fill_fine_same_level(U)

    # -- one Runge-Kutta step of one level (materialised fluxes, gamma-weighted sums) ---------------------------------------------
    def _level_step(self, L: _Level, dt: float, stage_fill):
        """advances L.S[L.cur] (ghosts filled by the caller); the new state becomes L.S[L.cur]; L.Facc / L.Sacc hold the
        flux / source sums of the step."""
        plan, a, b, g = L.plan, self.alpha, self.beta, self.gamma
        order = [L.cur] + [i for i in range(4) if i != L.cur]           # buffers of U^(0), U^(1), U^(2), U^(3)
        U = [L.S[i] for i in order]
        for t in L.Facc:
            t.zero_()
        L.Sacc.zero_()
        for sn in range(3):
            if sn > 0:
                stage_fill(U[sn], U[0])
            L.src.zero_()                                                # Euler::computeFluxesAndSourcesOnPatch zero-fills the source
            plan.compute_flux_and_source(U[sn], dt, L.F, L.src)
            none = [None] * sn
            gam = [0.0] * sn + [float(g[sn][sn])]
            plan.advance_stage(list(a[sn][:sn + 1]), list(b[sn][:sn + 1]), U[:sn + 1], none + [L.F], none + [L.src], U[sn + 1],
                               gamma=gam, F_acc=L.Facc, S_acc=L.Sacc)
        L.cur = order[3]

    # -- one coarse time step of the hierarchy ---------------------------------------------------------------------------------------
    def advance(self, dt: float):
        C, F, r = self.coarse, self.fine, self.r
        i_old = C.cur
        self._fill_coarse(C.S[i_old])
        self._level_step(C, dt, self._coarse_stage_fill)
        Uc_old, Uc_new = C.S[i_old], C.S[C.cur]
        self._fill_coarse(Uc_new)
        for t in self.fluxsum:
            t.zero_()                                                    # preprocessFluxAndSourceData, first fine step
        for s in range(r):
            self._fill_fine_from_coarse(F.S[F.cur], Uc_old, Uc_new, s / r)
            self._level_step(F, dt / r, self._fine_stage_fill)
            self.pair.fluxsum_update(F.Facc, self.fluxsum)
        # synchronizeLevelWithCoarser
        self.pair.coarsen_fluxsum(self.fluxsum, C.Facc)
        C.plan.advance_stage([1.0], [1.0], [Uc_old], [C.Facc], [C.Sacc], Uc_new)      # Euler::synchronizeFluxes on the old data
        self.pair.coarsen(F.S[F.cur], self.clo, self.chi, Uc_new)
        self.time += dt

    # -- diagnostics --------------------------------------------------------------------------------------------------------------------
    def composite_totals(self):
        """sum over the composite grid of every conserved component times the cell volume."""
        torch, dim = self.torch, self.dim
        dVc, dVf = float(np.prod(self.dxc)), float(np.prod(self.dxf))
        inner = self.Uc[self._inner()].clone()
        inner[(slice(None),) + tuple(slice(self.clo[a], self.chi[a]) for a in reversed(range(dim)))] = 0.0
        axes = tuple(range(1, dim + 1))
        return (inner.sum(dim=axes) * dVc + self.Uf[self._inner()].sum(dim=axes) * dVf).cpu().numpy()
