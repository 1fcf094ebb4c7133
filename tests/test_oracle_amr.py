"""SURVEY row f3, AMR plumbing on the CPU.

(1) The oracle's numpy restatements of the SAMRAI operators (oracle/amr.py; SAMRAI is an un-vendored dependency: parity
    unpinned) have the operators' defining properties: conservative refine reproduces linear data exactly,
    conserves the coarse cell's content and creates no new extrema; conservative coarsen is its left inverse on the means.
(2) The product's thread functions (hb2_amr.cuh, compiled by g++ in tests/host_emu/emu_amr.cpp) give the oracle's results
    BIT FOR BIT -- refine with time interpolation (ghost slabs reaching into the coarse ghosts), coarsen, flux integrals,
    their coarsening onto the coarse flux, FLOW extrapolation.
(3) A two-level step of the oracle hierarchy conserves every component on the composite grid to round-off (the flux
    correction closes the coarse-fine interface), keeps a uniform state, and a hierarchy whose fine patch covers the whole
    domain reproduces the single-level fine run on the fine level."""
import numpy as np
import pytest

import emu_host
from hamers_b200 import problems as pb
from oracle import amr
from oracle import oracle as orc

G = 4


def _geom(dim):
    nc = (12, 10, 8)[:dim]
    clo, chi = (3, 2, 1)[:dim], (9, 8, 6)[:dim]
    r = (2,) * dim
    dxc = (0.1, 0.2, 0.3)[:dim]
    dxf = tuple(h / 2 for h in dxc)
    nf = tuple(2 * (chi[a] - clo[a]) for a in range(dim))
    return nc, clo, chi, r, dxc, dxf, nf


@pytest.mark.parametrize("dim", [2, 3])
def test_refine_properties(dim):
    nc, clo, chi, r, dxc, dxf, nf = _geom(dim)
    shape = tuple(n + 2 * G for n in reversed(nc))
    xs = [(np.arange(-G, nc[a] + G) + 0.5) * dxc[a] for a in range(dim)]
    lin = sum(c * x.reshape([-1 if b == dim - 1 - a else 1 for b in range(dim)]) for a, (c, x) in enumerate(zip((1.5, -0.7, 0.3), xs))) + 2.0
    assert lin.shape == shape
    lo, hi = (-G,) * dim, tuple(n + G for n in nf)
    fine = amr.conservative_linear_refine(lin[None], dim, G, clo, r, dxc, dxf, lo, hi)[0]
    xf = [(clo[a] * dxc[a] + (np.arange(lo[a], hi[a]) + 0.5) * dxf[a]) for a in range(dim)]
    want = sum(c * x.reshape([-1 if b == dim - 1 - a else 1 for b in range(dim)]) for a, (c, x) in enumerate(zip((1.5, -0.7, 0.3), xf))) + 2.0
    assert np.allclose(fine, want, rtol=0, atol=1e-13)                       # exact on linear data
    rng = np.random.default_rng(1)
    Uc = rng.uniform(0.5, 2.0, (3,) + shape)
    lo, hi = (0,) * dim, nf
    fine = amr.conservative_linear_refine(Uc, dim, G, clo, r, dxc, dxf, lo, hi)
    back = amr.conservative_coarsen(fine, dim, r, dxc, dxf)
    inner = Uc[amr._sl(dim, clo, chi, G)]
    assert np.abs(back - inner).max() <= 4e-16 * np.abs(inner).max()        # conservative: the mean of a coarse cell is kept
    # the limiter's bound: per direction the correction is at most half the smaller one-sided difference (delta = dx_c/4,
    # slope <= 2 min(|dL|, |dR|)/dx_c), and zero at an extremum of that direction
    dev = np.zeros_like(inner)
    for a in range(dim):
        sh_p = [clo[b] + (1 if b == a else 0) for b in range(dim)]
        sh_m = [clo[b] - (1 if b == a else 0) for b in range(dim)]
        ext = [chi[b] - clo[b] for b in range(dim)]
        vp = Uc[amr._sl(dim, sh_p, [sh_p[b] + ext[b] for b in range(dim)], G)]
        vm = Uc[amr._sl(dim, sh_m, [sh_m[b] + ext[b] for b in range(dim)], G)]
        dR, dL = vp - inner, inner - vm
        dev += np.where(dL * dR > 0.0, 0.5 * np.minimum(np.abs(dR), np.abs(dL)), 0.0)
    rep = inner
    dev_f = dev
    for ax in range(1, dim + 1):
        rep, dev_f = np.repeat(rep, 2, axis=ax), np.repeat(dev_f, 2, axis=ax)
    assert (np.abs(fine - rep) <= dev_f + 1e-14).all()


@pytest.mark.parametrize("dim", [2, 3])
def test_emulated_operator_kernels_match_the_oracle_bit_for_bit(dim):
    nc, clo, chi, r, dxc, dxf, nf = _geom(dim)
    ncomp = neq = dim + 2
    rng = np.random.default_rng(7)
    cshape = (ncomp,) + tuple(n + 2 * G for n in reversed(nc))
    fshape = (ncomp,) + tuple(n + 2 * G for n in reversed(nf))
    Uold, Unew = rng.standard_normal(cshape), rng.standard_normal(cshape)
    p = emu_host.amr_pair(dim, nc, nf, r, clo, dxc, dxf, ncomp, neq)
    # refine into ghost slabs and into the whole ghost box, with and without time interpolation, both thread orders
    boxes = [((-G,) * dim, tuple(n + G for n in nf)), ((-G,) + (0,) * (dim - 1), (0,) + tuple(nf[1:]))]
    for lo, hi in boxes:
        for new, tfrac in ((Unew, 0.5), (Unew, 0.25), (None, 0.0)):
            want = amr.conservative_linear_refine(Uold if new is None else amr.time_interpolate(Uold, new, tfrac), dim, G, clo, r,
                                                  dxc, dxf, lo, hi)
            for rev in (False, True):
                Uf = np.full(fshape, np.nan)
                emu_host.amr_refine(p, Uold, new, tfrac, lo, hi, Uf, reverse=rev)
                assert np.array_equal(Uf[amr._sl(dim, lo, hi, G)], want)
                Uf[amr._sl(dim, lo, hi, G)] = np.nan
                assert np.isnan(Uf).all()                    # nothing outside the box is written
    # coarsen
    Uf = rng.standard_normal(fshape)
    Uc = np.full(cshape, np.nan)
    emu_host.amr_coarsen(p, Uf, clo, chi, Uc)
    want = amr.conservative_coarsen(Uf[amr._sl(dim, (0,) * dim, nf, G)], dim, r, dxc, dxf)
    assert np.array_equal(Uc[amr._sl(dim, clo, chi, G)], want)
    # flux integrals over two fine steps, then onto the coarse flux
    fdesc = orc.PatchDesc(dim=dim, n=nf)
    cdesc = orc.PatchDesc(dim=dim, n=nc)
    fsum = [np.zeros((neq,) + tuple(nf[a] for a in reversed(range(dim)) if a != d)) for d in range(dim) for _ in (0, 1)]
    want_sum = [None] * (2 * dim)
    for step in range(2):
        F = [rng.standard_normal((neq,) + fdesc.side_shape(d)) for d in range(dim)]
        emu_host.amr_fluxsum(p, F, fsum)
        for d in range(dim):
            for side in (0, 1):
                o = amr.outer_side(F[d], dim, d, side)
                want_sum[2 * d + side] = 0.0 + o if step == 0 else want_sum[2 * d + side] + o
    for k in range(2 * dim):
        assert np.array_equal(fsum[k], want_sum[k])
    Fc = [rng.standard_normal((neq,) + cdesc.side_shape(d)) for d in range(dim)]
    Fc_want = [f.copy() for f in Fc]
    emu_host.amr_coarsen_fluxsum(p, fsum, Fc)
    for d in range(dim):
        for side in (0, 1):
            idx = [slice(None)] + [slice(clo[a], chi[a]) for a in reversed(range(dim))]
            idx[dim - d] = clo[d] if side == 0 else chi[d]
            Fc_want[d][tuple(idx)] = amr.coarsen_outer_side(want_sum[2 * d + side], dim, d, r, dxc, dxf)
        assert np.array_equal(Fc[d], Fc_want[d])
    # FLOW extrapolation
    for d in range(dim):
        for side in (0, 1):
            U = rng.standard_normal(cshape)
            W = U.copy()
            emu_host.amr_extrapolate(dim, nc, G, U, d, side)
            amr.fill_extrapolate(W, dim, G, d, side)
            assert np.array_equal(U, W) and not np.array_equal(U, rng.standard_normal(cshape))


def _hierarchy(desc, clo, chi, periodic, U, fine_ic):
    H = amr.TwoLevelOracle(desc, clo, chi, 2, periodic)
    dim = desc.dim
    H.Uc[amr._sl(dim, (0,) * dim, desc.n, G)] = U
    H.Uf[amr._sl(dim, (0,) * dim, H.df.n, G)] = fine_ic[(slice(None),) + tuple(slice(2 * clo[a], 2 * chi[a]) for a in reversed(range(dim)))]
    H.Uc[amr._sl(dim, clo, chi, G)] = amr.conservative_coarsen(H.Uf[amr._sl(dim, (0,) * dim, H.df.n, G)], dim, H.r, desc.dx, H.df.dx)
    return H


@pytest.mark.parametrize("model", [0, 2])
@pytest.mark.parametrize("clo,chi", [((8, 8), (24, 24)), ((8, 0), (24, 32)), ((0, 4), (32, 20))])
def test_two_level_step_conserves_on_the_composite_grid(clo, chi, model, oracle_lib):
    N = 32
    if model == 0:
        U, dx, gam = pb.convergence_single_species(2, N)
        Uf, _, _ = pb.convergence_single_species(2, 2 * N)
        desc = oracle_lib.PatchDesc(dim=2, n=(N, N), model=0, ns=1, gamma=gam, dx=dx)
    else:
        U, dx, gam, R = pb.convergence_four_eqn(2, N)
        Uf, _, _, _ = pb.convergence_four_eqn(2, 2 * N)
        desc = oracle_lib.PatchDesc(dim=2, n=(N, N), model=2, ns=2, gamma=gam, R=R, dx=dx, scheme=2)
    H = _hierarchy(desc, clo, chi, (True, True), U, Uf)
    t0 = H.composite_totals()
    for _ in range(3):
        H.advance(0.2 * dx[0])
    t1 = H.composite_totals()
    assert np.isfinite(H.Uc[amr._sl(2, (0, 0), desc.n, G)]).all()
    assert (np.abs(t1 - t0) <= 4e-15 * np.abs(t0)).all(), (t0, t1)        # total mass, momentum, energy to round-off
    # without the flux correction the interface does NOT close: the check above is not vacuous
    H2 = _hierarchy(desc, clo, chi, (True, True), U, Uf)
    H2.r_backup = amr.coarsen_outer_side
    try:
        calls = []

        def skewed(fsum, dim, d, r, dxc, dxf):        # every interface side is off by a different fraction of a per cent
            calls.append(d)
            return (1.0 - 0.01 / len(calls)) * H2.r_backup(fsum, dim, d, r, dxc, dxf)

        amr.coarsen_outer_side = skewed
        H2.advance(0.2 * dx[0])
    finally:
        amr.coarsen_outer_side = H2.r_backup
    assert (np.abs(H2.composite_totals() - t0) > 2e-13 * np.abs(t0)).any()


def test_uniform_state_is_preserved_and_full_cover_equals_the_fine_level(oracle_lib):
    N = 16
    U, dx, gam = pb.convergence_single_species(2, N)
    desc = oracle_lib.PatchDesc(dim=2, n=(N, N), model=0, ns=1, gamma=gam, dx=dx)
    const = np.broadcast_to(np.array([1.3, 0.4, -0.2, 3.0])[:, None, None], U.shape).copy()
    H = _hierarchy(desc, (4, 4), (12, 12), (True, False), const, np.broadcast_to(const[:, :1, :1], (4, 2 * N, 2 * N)).copy())
    H.advance(0.1 * dx[0])
    assert np.abs(H.Uc[amr._sl(2, (0, 0), desc.n, G)] - const).max() < 1e-14
    assert np.abs(H.Uf[amr._sl(2, (0, 0), H.df.n, G)] - const[:, :1, :1]).max() < 1e-14
    # fine patch over the whole periodic domain: the fine level never sees the coarse one
    Uf, dxf, _ = pb.convergence_single_species(2, 2 * N)
    H = _hierarchy(desc, (0, 0), (N, N), (True, True), U, Uf)
    dt = 0.2 * dx[0]
    H.advance(dt)
    fine = oracle_lib.PatchDesc(dim=2, n=(2 * N, 2 * N), model=0, ns=1, gamma=gam, dx=dxf)
    ref = Uf.copy()
    oracle_lib.level_advance(fine, (8, 8), ref, dt / 2, 2, nthreads=0)
    got = H.Uf[amr._sl(2, (0, 0), H.df.n, G)]
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()
