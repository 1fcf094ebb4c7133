"""The two-level hierarchy driver of hamers_b200/amr.py on the CPU: abi.Plan / abi.AmrPair are replaced by
stand-ins that run the host emulation of the very thread functions (tests/emu_host.py; the RK update from materialised fluxes
through the oracle's advance_stage, its gamma-weighted sums in numpy) on CPU tensors.  Coarse steps of the product driver
are compared BIT FOR BIT with the oracle's independent restatement of the same sequence (oracle/amr.py): which buffer holds
which stage state, which ghost slabs come from the coarser level and when, the order of the synchronisation.  The GPU
counterpart is tests/test_zz_gpu_amr.py."""
import numpy as np
import pytest

import emu_host
from hamers_b200 import abi
from hamers_b200 import problems as pb
from hamers_b200.amr import TwoLevelHierarchy
from oracle import amr
from oracle import oracle as orc

G = 4


class EmuPlan:
    def __init__(self, dim, n, flow_model=0, species_gamma=(1.4,), species_R=(), dx=(1.0, 1.0, 1.0), math=0, scheme=0, **kw):
        ns = 1 if flow_model == 0 else len(species_gamma)
        self.desc = orc.PatchDesc(dim=dim, n=tuple(n), model=flow_model, ns=ns, gamma=tuple(species_gamma), R=tuple(species_R),
                                  dx=tuple(dx), scheme=scheme)
        self.dim, self.neq, self.ncomp, self.math = dim, self.desc.neq, self.desc.ncomp, math
        self.ghost_shape, self.cell_shape = self.desc.ghost_shape, self.desc.cell_shape

    def side_shape(self, d):
        return self.desc.side_shape(d)

    def use_torch_stream(self):
        return self

    def close(self):
        pass

    def compute_flux_and_source(self, Q, dt, flux, source=None):
        F, _ = emu_host.flux_and_source(self.desc, Q.numpy(), dt, math=self.math, source=source.numpy() if source is not None else None)
        for a, f in enumerate(F):
            flux[a].numpy()[...] = f

    def advance_stage(self, alpha, beta, U_int, F_int, S_int, U_out, gamma=None, F_acc=None, S_acc=None):
        Fl = [None if f is None else [x.numpy() for x in f] for f in F_int]
        Sl = [None if s is None else s.numpy() for s in S_int]
        out = orc.advance_stage(self.desc, list(alpha), list(beta), [u.numpy() for u in U_int], Fl, Sl)
        inner = (slice(None),) + (slice(G, -G),) * self.dim
        U_out.numpy()[inner] = out[inner]                # the kernel leaves the ghosts of U_out alone
        if gamma is not None:
            for m, gm in enumerate(gamma):
                if gm != 0.0 and F_int[m] is not None:
                    for d in range(self.dim):
                        F_acc[d].numpy()[...] = F_acc[d].numpy() + gm * F_int[m][d].numpy()
                    S_acc.numpy()[...] = S_acc.numpy() + gm * S_int[m].numpy()

    def fill_ghosts_periodic(self, U, mask=7):
        amr.fill_periodic(U.numpy(), self.dim, G, [a for a in range(self.dim) if (mask >> a) & 1])

    def fill_ghosts_extrapolate(self, U, direction, side):
        emu_host.amr_extrapolate(self.dim, self.desc.n, G, U.numpy(), direction, side)


class EmuPair:
    def __init__(self, dim, nc, nf, ratio, origin, dxc, dxf, ncomp, neq):
        self.p = emu_host.amr_pair(dim, nc, nf, ratio, origin, dxc, dxf, ncomp, neq)

    def refine(self, Uc_old, Uc_new, tfrac, lo, hi, Uf):
        emu_host.amr_refine(self.p, Uc_old.numpy(), None if Uc_new is None else Uc_new.numpy(), tfrac, lo, hi, Uf.numpy())

    def coarsen(self, Uf, lo, hi, Uc):
        emu_host.amr_coarsen(self.p, Uf.numpy(), lo, hi, Uc.numpy())

    def fluxsum_update(self, F, fsum):
        emu_host.amr_fluxsum(self.p, [f.numpy() for f in F], [s.numpy() for s in fsum])

    def coarsen_fluxsum(self, fsum, Fc):
        emu_host.amr_coarsen_fluxsum(self.p, [s.numpy() for s in fsum], [f.numpy() for f in Fc])


def build_pair(model, N, clo, chi, periodic, factory_kw, scheme=0):
    """(product hierarchy, oracle hierarchy) with the same initial data"""
    if model == 0:
        U, dx, gam = pb.random_state(2, (N, N), model=0, seed=5, shock=False)
        Uf, _, _ = pb.random_state(2, (2 * N, 2 * N), model=0, seed=6, shock=False)
        R = ()
        desc = orc.PatchDesc(dim=2, n=(N, N), model=0, ns=1, gamma=gam, dx=dx, scheme=scheme)
    else:
        U, dx, gam, R = pb.convergence_four_eqn(2, N)
        Uf, _, _, _ = pb.convergence_four_eqn(2, 2 * N)
        desc = orc.PatchDesc(dim=2, n=(N, N), model=2, ns=2, gamma=gam, R=R, dx=dx, scheme=scheme)
    # tame the white noise so that a few steps stay well inside the physical range
    if model == 0:
        U = 0.9 * pb.convergence_single_species(2, N)[0] + 0.1 * U
        Uf = 0.9 * pb.convergence_single_species(2, 2 * N)[0] + 0.1 * Uf
    O = amr.TwoLevelOracle(desc, clo, chi, 2, periodic)
    H = TwoLevelHierarchy(2, (N, N), clo, chi, 2, periodic, flow_model=model, species_gamma=gam, species_R=R, dx=dx,
                          math=abi.MATH_EXACT, scheme=scheme, **factory_kw)
    fine_box = (slice(None),) + tuple(slice(2 * clo[a], 2 * chi[a]) for a in (1, 0))
    inner = (slice(None),) + (slice(G, -G),) * 2
    O.Uc[inner] = U
    O.Uf[inner] = Uf[fine_box]
    O.Uc[amr._sl(2, clo, chi, G)] = amr.conservative_coarsen(O.Uf[inner], 2, O.r, desc.dx, O.df.dx)
    H.set_coarse(U)
    H.set_fine(np.ascontiguousarray(Uf[fine_box]))
    H.coarsen_fine_onto_coarse()
    return H, O


@pytest.mark.parametrize("model,scheme", [(0, 0), (2, 2)])
@pytest.mark.parametrize("clo,chi,periodic", [((4, 4), (12, 12), (True, True)), ((4, 0), (12, 16), (False, True)),
                                             ((0, 0), (10, 16), (False, True))])
def test_product_driver_over_emulated_kernels_equals_the_oracle_hierarchy(clo, chi, periodic, model, scheme, oracle_lib):
    N = 16
    H, O = build_pair(model, N, clo, chi, periodic, dict(device="cpu", plan_factory=EmuPlan, pair_factory=EmuPair), scheme)
    inner = (slice(None),) + (slice(G, -G),) * 2
    assert np.array_equal(H.Uc.numpy()[inner], O.Uc[inner])
    t0 = H.composite_totals()
    dt = 0.1 * O.dc.dx[0]
    for _ in range(2):
        H.advance(dt)
        O.advance(dt)
        assert np.isfinite(O.Uc[inner]).all()
        assert np.array_equal(H.Uc.numpy()[inner], O.Uc[inner])
        assert np.array_equal(H.Uf.numpy()[inner], O.Uf[inner])
    if all(periodic):
        assert (np.abs(H.composite_totals() - t0) <= 4e-15 * np.abs(t0)).all()
    assert np.array_equal(H.composite_totals(), O.composite_totals()) or np.allclose(H.composite_totals(), O.composite_totals(), rtol=1e-14)
