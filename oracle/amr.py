"""CPU ORACLE (test infrastructure, NOT a product path) of the two-level AMR step around the convective hot path
(SURVEY row f3): numpy restatements of the patch-data operators and of the level / synchronisation sequence, composed with
the C oracle's flux and stage functions (oracle/hamers_oracle.c).

What is restated, and from where (path:line under the reference tree):
  * the sequence: RungeKuttaLevelIntegrator::advanceLevel (src/algs/integrator/RungeKuttaLevelIntegrator.cpp:1457-1929: ghost
    fill at the current time, zeroed flux / source sums, per stage a same-level fill that leaves the coarse-fine ghost values
    of the first fill in place ("Dirichlet", :1672-1745), gamma-weighted flux sums (Euler.cpp:1555-1640)),
    postprocessFluxAndSourceData (:2968-3230; algs_upfluxsum2d.f / 3d.f: fluxsum += flux on the outer sides),
    synchronizeLevelWithCoarser (:2131-2209: coarsen the flux integrals onto the coarse flux, repeat the conservative
    difference Euler::synchronizeFluxes, Euler.cpp:1682-1949, coarsen the fine solution) and SAMRAI's
    TimeRefinementIntegrator order (coarse step first, then `ratio` fine steps, then the synchronisation);
  * BDRY_COND::BASIC::FLOW: src/util/basic_boundary_conditions/BasicCartesianBoundaryUtilities2.cpp:310-345;
  * the three SAMRAI geom / pdat operators named at FlowModelSingleSpecies.cpp:578-579 and Euler.cpp:307
    ("CONSERVATIVE_LINEAR_REFINE", "CONSERVATIVE_COARSEN", linear time interpolation).  SAMRAI (v4.1.0,
    circleci/install-SAMRAI.sh:4) is an un-vendored dependency: its algorithm is restated here from its published source
    (geom_cartrefine{2,3}d.m4 cartclinrefcelldoub*, geom_cartcoarsen{2,3}d.m4 cartwgtavgcelldoub* / cartwgtavgoutsiddoub*,
    pdat lintimeintcelldoub*).  PARITY UNPINNED for these (the reference holds no multi-level test):
    tests/test_oracle_amr.py checks their defining properties instead.

Arrays: (ncomp, [z,] y, x), x fastest; cell data carry g ghost cells (the SAMRAI ghost box), side data none."""
from __future__ import annotations

import numpy as np

from . import oracle as orc

G = 4


def _sl(dim, lo, hi, g):
    """numpy slices of the cells [lo, hi) (patch indices, x first) of a ghost-box array with g ghosts."""
    return (slice(None),) + tuple(slice(lo[a] + g, hi[a] + g) for a in reversed(range(dim)))


def time_interpolate(old, new, tfrac):
    """pdat CellDoubleLinearTimeInterpolateOp: arrayold*oldfrac + arraynew*tfrac."""
    oldfrac = 1.0 - tfrac
    return old * oldfrac + new * tfrac


def conservative_linear_refine(Uc, dim, gc, origin, ratio, dxc, dxf, lo, hi):
    """geom CartesianCellDoubleConservativeLinearRefine on the fine cells [lo, hi) (fine patch indices).  Uc: coarse
    ghost-box array; origin: coarse-patch index of the coarse cell that holds fine cell 0.  Per direction the slope is the
    central difference limited to twice the smaller one-sided difference, zero where the differences change sign."""
    idx, delta = [], []
    for a in range(dim):
        f = np.arange(lo[a], hi[a])
        q = np.floor_divide(f, ratio[a])
        ir = f - q * ratio[a]
        idx.append(q + origin[a] + gc)
        delta.append((ir.astype(np.float64) + 0.5) * dxf[a] - dxc[a] * 0.5)

    def take(shift_axis=None, shift=0):
        ix = [idx[a] + (shift if a == shift_axis else 0) for a in range(dim)]
        if dim == 2:
            return Uc[:, ix[1][:, None], ix[0][None, :]]
        return Uc[:, ix[2][:, None, None], ix[1][None, :, None], ix[0][None, None, :]]

    v0 = take()
    val = v0.copy()
    for a in range(dim):
        dR = take(a, 1) - v0
        dL = v0 - take(a, -1)
        coef2 = 0.5 * (dR + dL)
        bound = 2.0 * np.minimum(np.abs(dR), np.abs(dL))
        slope = np.where(dL * dR > 0.0, np.copysign(np.minimum(np.abs(coef2), bound), coef2) / dxc[a], 0.0)
        shape = [1] * (dim + 1)
        shape[dim - a] = -1
        val = val + slope * delta[a].reshape(shape)
    return val


def conservative_coarsen(Uf_box, dim, ratio, dxc, dxf):
    """geom CartesianCellDoubleWeightedAverage: Uf_box holds the fine cells of whole coarse cells, (ncomp, [z,] y, x);
    sum of the fine values times dV_f in the order ir2, ir1, ir0 (ir0 fastest), divided by dV_c."""
    dVf, dVc = dxf[0] * dxf[1], dxc[0] * dxc[1]
    if dim == 3:
        dVf, dVc = dVf * dxf[2], dVc * dxc[2]
    r = list(ratio[:dim]) + [1] * (3 - dim)
    spv = None
    for i2 in range(r[2]):
        for i1 in range(r[1]):
            for i0 in range(r[0]):
                if dim == 2:
                    part = Uf_box[:, i1::r[1], i0::r[0]] * dVf
                else:
                    part = Uf_box[:, i2::r[2], i1::r[1], i0::r[0]] * dVf
                spv = 0.0 + part if spv is None else spv + part
    return spv / dVc


def outer_side(F, dim, direction, side):
    """The faces of a ghost-0 side array (neq, [z,] y, x) on side 0 / 1 of the patch: (neq, tangential cells...)."""
    ax = dim - direction          # numpy axis of the direction (axis 0 is the component)
    return np.take(F, 0 if side == 0 else F.shape[ax] - 1, axis=ax)


def coarsen_outer_side(fsum, dim, direction, ratio, dxc, dxf):
    """geom CartesianOutersideDoubleWeightedAverage: fsum (neq, tangential fine cells, lower direction fastest) ->
    (neq, tangential coarse cells): sum times the fine face area, divided by the coarse face area."""
    tang = [a for a in range(dim) if a != direction]
    areaf, areac = dxf[tang[0]], dxc[tang[0]]
    if dim == 3:
        areaf, areac = areaf * dxf[tang[1]], areac * dxc[tang[1]]
    ra = ratio[tang[0]]
    rb = ratio[tang[1]] if dim == 3 else 1
    spv = None
    for ib in range(rb):
        for ia in range(ra):
            part = (fsum[:, ia::ra] if dim == 2 else fsum[:, ib::rb, ia::ra]) * areaf
            spv = 0.0 + part if spv is None else spv + part
    return spv / areac


def fill_periodic(U, dim, g, axes):
    """same-level periodic fill of the directions in `axes`, one after the other, ghost-inclusive in the others."""
    for a in axes:
        ax = dim - a
        n = U.shape[ax] - 2 * g
        src_lo = [slice(None)] * U.ndim
        dst_lo = [slice(None)] * U.ndim
        src_hi = [slice(None)] * U.ndim
        dst_hi = [slice(None)] * U.ndim
        idx = (np.arange(-g, 0) % n) + g
        dst_lo[ax] = slice(0, g)
        U[tuple(dst_lo)] = np.take(U, idx, axis=ax)
        idx = (np.arange(n, n + g) % n) + g
        dst_hi[ax] = slice(n + g, n + 2 * g)
        U[tuple(dst_hi)] = np.take(U, idx, axis=ax)
        del src_lo, src_hi


def fill_extrapolate(U, dim, g, direction, side):
    """BDRY_COND::BASIC::FLOW on face (direction, side): ghost cells copy the adjacent interior cell; the other directions
    run over the interior."""
    ax = dim - direction
    n = U.shape[ax] - 2 * g
    inner = [slice(None)] + [slice(g, -g)] * dim
    dst, src = list(inner), list(inner)
    if side == 0:
        dst[ax], src[ax] = slice(0, g), slice(g, g + 1)
    else:
        dst[ax], src[ax] = slice(n + g, n + 2 * g), slice(n + g - 1, n + g)
    U[tuple(dst)] = U[tuple(src)]


class TwoLevelOracle:
    """One coarse patch covering the domain and one fine patch refined by `ratio` over the coarse cells [clo, chi).
    periodic[a]: the domain is periodic in direction a, otherwise both faces are BDRY_COND::BASIC::FLOW."""

    def __init__(self, desc_c: orc.PatchDesc, clo, chi, ratio=2, periodic=(True, True, True)):
        dim = desc_c.dim
        self.dim, self.r = dim, [int(ratio)] * dim
        self.clo, self.chi = tuple(clo), tuple(chi)
        self.periodic = tuple(periodic[:dim])
        self.dc = desc_c
        nf = tuple(self.r[a] * (chi[a] - clo[a]) for a in range(dim))
        dxf = tuple(desc_c.dx[a] / self.r[a] for a in range(dim))
        import dataclasses

        self.df = dataclasses.replace(desc_c, n=nf, dx=dxf)
        # a direction the fine patch spans completely AND that is periodic is filled from the fine patch itself
        self.self_periodic = tuple(self.periodic[a] and clo[a] == 0 and chi[a] == desc_c.n[a] for a in range(dim))
        # a fine face on a non-periodic domain boundary is a physical boundary of the fine level too
        self.fine_physical = tuple((not self.periodic[a] and clo[a] == 0, not self.periodic[a] and chi[a] == desc_c.n[a])
                                   for a in range(dim))
        self.Uc = np.zeros((desc_c.ncomp,) + desc_c.ghost_shape)
        self.Uf = np.zeros((self.df.ncomp,) + self.df.ghost_shape)

    # -- ghost fills ------------------------------------------------------------------------------------------------
    def fill_coarse(self, U):
        for a in range(self.dim):
            if not self.periodic[a]:
                fill_extrapolate(U, self.dim, G, a, 0)
                fill_extrapolate(U, self.dim, G, a, 1)
        fill_periodic(U, self.dim, G, [a for a in range(self.dim) if self.periodic[a]])
        # non-periodic directions: the corner ghosts come from the periodic fill of the extrapolated slabs; where two
        # non-periodic directions meet they stay unfilled (the path never reads corner ghosts)

    def cf_boxes(self):
        """Fine ghost slabs filled from the coarser level: for every direction that is not self-periodic both ghost slabs,
        interior extent in the self-periodic directions (their ghosts follow by the periodic fill), ghost-inclusive in the
        other directions.  Physical-boundary slabs are excluded."""
        dim, nf = self.dim, self.df.n
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

    def fill_fine_same_level(self, U):
        for a in range(self.dim):
            for side in (0, 1):
                if self.fine_physical[a][side]:
                    fill_extrapolate(U, self.dim, G, a, side)
        fill_periodic(U, self.dim, G, [a for a in range(self.dim) if self.self_periodic[a]])

    def fill_fine_from_coarse(self, U, Uc_old, Uc_new, tfrac):
        Uc = Uc_old if Uc_new is None else time_interpolate(Uc_old, Uc_new, tfrac)
        for lo, hi in self.cf_boxes():
            U[_sl(self.dim, lo, hi, G)] = conservative_linear_refine(Uc, self.dim, G, self.clo, self.r, self.dc.dx, self.df.dx, lo, hi)
        self.fill_fine_same_level(U)

    def initialize_fine_from_coarse(self):
        """fill the whole fine patch from the coarse data (what a new level gets before the user IC overwrites it)."""
        self.fill_coarse(self.Uc)
        lo, hi = (0,) * self.dim, self.df.n
        self.Uf[_sl(self.dim, lo, hi, G)] = conservative_linear_refine(self.Uc, self.dim, G, self.clo, self.r, self.dc.dx,
                                                                     self.df.dx, lo, hi)

    # -- one Runge-Kutta step of one level with the materialised fluxes and their gamma-weighted sums -----------------------
    def level_step(self, desc, U0, dt, stage_fill):
        a, b, g = orc.SSPRK3_ALPHA, orc.SSPRK3_BETA, orc.SSPRK3_GAMMA
        Uint, Fint, Sint = [U0], [], []
        Facc = [np.zeros((desc.neq,) + desc.side_shape(d)) for d in range(desc.dim)]
        Sacc = np.zeros((desc.neq,) + desc.cell_shape)
        for sn in range(3):
            if sn > 0:
                stage_fill(Uint[sn], U0)
            F, S = orc.compute_flux_and_source(desc, Uint[sn], dt)
            Fint.append(F)
            Sint.append(S)
            Fl = [Fint[m] if b[sn][m] != 0.0 else None for m in range(sn + 1)]
            Sl = [Sint[m] if b[sn][m] != 0.0 else None for m in range(sn + 1)]
            Unew = orc.advance_stage(desc, list(a[sn][:sn + 1]), list(b[sn][:sn + 1]), Uint[:sn + 1], Fl, Sl)
            for d in range(desc.dim):
                Facc[d] = Facc[d] + g[sn][sn] * F[d]
            Sacc = Sacc + g[sn][sn] * S
            Uint.append(Unew)
        return Uint[3], Facc, Sacc

    def _coarse_stage_fill(self, U, U0):
        self.fill_coarse(U)

    def _fine_stage_fill(self, U, U0):
        for lo, hi in self.cf_boxes():
            s = _sl(self.dim, lo, hi, G)
            U[s] = U0[s]                    # the coarse-fine ghost values of the first fill stay in place
        self.fill_fine_same_level(U)

    # -- one coarse time step of the hierarchy -----------------------------------------------------------------------------
    def advance(self, dt):
        dim, r = self.dim, self.r[0]
        dc, df = self.dc, self.df
        Uc_old = self.Uc
        self.fill_coarse(Uc_old)
        Uc_new, Facc_c, Sacc_c = self.level_step(dc, Uc_old, dt, self._coarse_stage_fill)
        self.fill_coarse(Uc_new)
        fsum = {}
        for s in range(r):
            self.fill_fine_from_coarse(self.Uf, Uc_old, Uc_new, s / r)
            Uf_new, Facc_f, _ = self.level_step(df, self.Uf, dt / r, self._fine_stage_fill)
            for d in range(dim):
                for side in (0, 1):
                    o = outer_side(Facc_f[d], dim, d, side)
                    fsum[(d, side)] = (0.0 + o) if s == 0 else fsum[(d, side)] + o
            self.Uf = Uf_new
        # synchronizeLevelWithCoarser: flux integrals -> coarse flux on the fine patch's boundary
        for d in range(dim):
            for side in (0, 1):
                cf = coarsen_outer_side(fsum[(d, side)], dim, d, self.r, dc.dx, df.dx)
                face = self.clo[d] if side == 0 else self.chi[d]
                idx = [slice(None)] + [slice(self.clo[a], self.chi[a]) for a in reversed(range(dim))]
                idx[dim - d] = face
                Facc_c[d][tuple(idx)] = cf
        # repeat the conservative difference on the coarse level from the old data (Euler::synchronizeFluxes)
        Uc_sync = orc.advance_stage(dc, [1.0], [1.0], [Uc_old], [Facc_c], [Sacc_c])
        # conservative coarsen of the fine solution
        inner_f = _sl(dim, (0,) * dim, df.n, G)
        Uc_sync[_sl(dim, self.clo, self.chi, G)] = conservative_coarsen(self.Uf[inner_f], dim, self.r, dc.dx, df.dx)
        self.Uc = Uc_sync

    # -- diagnostics ----------------------------------------------------------------------------------------------------------
    def composite_totals(self):
        """sum over the composite grid of every conserved component times the cell volume (coarse cells under the fine
        patch replaced by their fine cells)."""
        dim = self.dim
        dVc, dVf = float(np.prod(self.dc.dx[:dim])), float(np.prod(self.df.dx[:dim]))
        inner_c = self.Uc[_sl(dim, (0,) * dim, self.dc.n, G)].copy()
        inner_c[(slice(None),) + tuple(slice(self.clo[a], self.chi[a]) for a in reversed(range(dim)))] = 0.0
        axes = tuple(range(1, dim + 1))
        return inner_c.sum(axis=axes) * dVc + self.Uf[_sl(dim, (0,) * dim, self.df.n, G)].sum(axis=axes) * dVf
