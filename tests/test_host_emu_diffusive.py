"""Host emulation of the diffusive-flux kernels against the oracle: the per-thread functions of
hamers_b200/csrc/hb2_diffusive.cuh, compiled with g++ and driven by plain loops, must be bit-identical to
oracle/oracle_diffusive.c (reference operation order, no FMA contraction on either side)."""
import numpy as np
import pytest

import emu_host
from hamers_b200 import problems as pb
from oracle import oracle as orc

TR = orc.Transport(mu=0.05, mu_v=0.02, c_p=3.5, c_v=2.5, Pr=0.72)


def state(dim, N, seed=5):
    U, dx, gam = pb.random_state(dim, N, seed=seed, shock=True)
    return orc.PatchDesc(dim=dim, n=N, gamma=gam, dx=dx), U


@pytest.mark.parametrize("dim,N", [(2, (24, 17)), (3, (13, 10, 12)), (2, (1, 1)), (3, (1, 2, 1)), (3, (7, 1, 3)), (2, (6, 40))])
def test_emulated_diffusive_flux_is_bit_identical(dim, N):
    desc, U = state(dim, N)
    Q = pb.pad_periodic(U, orc.GD)
    dt = 1.0e-3
    Fo = orc.compute_diffusive_flux(desc, TR, Q, dt)
    Fe = emu_host.diffusive_flux(desc, TR, Q, dt)
    for a in range(dim):
        assert not np.isnan(Fe[a]).any()
        assert np.array_equal(Fe[a], Fo[a]), f"dir {a}"
        assert not Fo[a][0].any() and not np.signbit(Fe[a][0]).any()        # +0.0 mass flux
        assert min(N) < 3 or np.abs(Fo[a][1:]).max() > 0.0


@pytest.mark.parametrize("dim", [2, 3])
def test_emulated_kernels_read_only_what_they_were_given(dim):
    """Ghost cells outside the stencil footprint (the corners beyond 3 cells in two directions at once are still needed:
    d/dy at x-ghost nodes) -- poisoning the layers the reference never reads must not change the result, poisoning one it
    reads must."""
    N = (9, 8, 7)[:dim]
    desc, U = state(dim, N)
    Q = pb.pad_periodic(U, orc.GD)
    F0 = emu_host.diffusive_flux(desc, TR, Q, 1.0e-3)
    Qp = Q.copy()
    g = orc.GD
    # a cell 4 deep in the x-ghosts AND 4 deep in the y-ghosts is outside every stencil (footprint: <= 3 in the second direction)
    idx = (slice(None),) + (slice(None),) * (dim - 2) + (g - 4, g - 4)
    Qp[idx] = 1.0e30
    F1 = emu_host.diffusive_flux(desc, TR, Qp, 1.0e-3)
    assert all(np.array_equal(a, b) for a, b in zip(F0, F1))
    Qp = Q.copy()
    idx = (1,) + (slice(None),) * (dim - 2) + (g - 3, g - 3)        # x momentum only: scaling every component keeps u and T
    Qp[idx] *= 1.5
    F2 = emu_host.diffusive_flux(desc, TR, Qp, 1.0e-3)
    assert any(not np.array_equal(a, b) for a, b in zip(F0, F2))


@pytest.mark.parametrize("dim,g", [(2, 6), (3, 6), (3, 4)])
def test_emulated_ns_stage_is_bit_identical(dim, g):
    rng = np.random.default_rng(11)
    N = (6, 5, 4)[:dim]
    desc = orc.PatchDesc(dim=dim, n=N, gamma=(1.4,), dx=(0.1, 0.2, 0.3)[:dim])
    neq = desc.neq
    shape = tuple(n + 2 * g for n in reversed(N))
    for alpha, beta in (([1.0], [1.0]), ([0.75, 0.25], [0.0, 0.25]), ([1.0 / 3.0, 0.0, 2.0 / 3.0], [0.0, 0.0, 2.0 / 3.0])):
        nc = len(alpha)
        U = [rng.standard_normal((neq,) + shape) for _ in range(nc)]
        Fc = [[rng.standard_normal((neq,) + desc.side_shape(a)) for a in range(dim)] for _ in range(nc)]
        Fd = [[rng.standard_normal((neq,) + desc.side_shape(a)) for a in range(dim)] for _ in range(nc)]
        S = [rng.standard_normal((neq,) + desc.cell_shape) for _ in range(nc)]
        Uo = orc.advance_stage_ns(desc, g, alpha, beta, U, Fc, Fd, S)
        Ue = emu_host.advance_stage_ns(desc, TR, g, alpha, beta, U, Fc, Fd, S)
        inner = (slice(None),) + (slice(g, -g),) * dim
        assert np.array_equal(Ue[inner], Uo[inner])
        Ue[inner] = np.nan
        assert np.isnan(Ue).all()                    # the kernel leaves the ghosts alone


@pytest.mark.parametrize("dim,N", [(2, (7, 5)), (3, (6, 4, 9)), (3, (2, 1, 3)), (2, (1, 1))])
def test_emulated_periodic_fill_and_ghost_view(dim, N):
    """Six-ghost periodic fill (also where the patch is narrower than the ghost width: images wrap several times) and the
    four-ghost view handed to the convective reconstructor, against numpy's wrap padding."""
    desc, U = state(dim, N)
    want6 = pb.pad_periodic(U, 6)
    got = np.full_like(want6, np.nan)
    got[(slice(None),) + (slice(6, -6),) * dim] = U
    emu_host.diff_fill_periodic(desc, TR, got)
    assert np.array_equal(got, want6)
    for g in (4, 0, 6):
        assert np.array_equal(emu_host.diff_extract_view(desc, TR, want6, g), pb.pad_periodic(U, g) if g else U)
    # mask: only x periodic -> y (and z) ghost layers keep their NaN, x ghosts of interior rows are filled
    part = np.full_like(want6, np.nan)
    inner = (slice(None),) + (slice(6, -6),) * dim
    part[inner] = U
    emu_host.diff_fill_periodic(desc, TR, part, mask=1)
    rows = (slice(None),) + (slice(6, -6),) * (dim - 1) + (slice(None),)
    assert np.array_equal(part[rows], want6[rows])
    part[rows] = 0.0
    assert np.isnan(part[:, 0]).all() and np.isnan(part[:, -1]).all()
    # across ranks (process grid (2,1,1)): the x ghosts of the interior rows arrive by the exchange, the rank-local
    # directions are filled afterwards and must be ghost-inclusive in x, so that edge and corner ghosts are filled too
    # (NavierStokesLevel.fill_ghosts: exchange, then fill_local(local_mask))
    exch = np.full_like(want6, np.nan)
    exch[rows] = want6[rows]
    emu_host.diff_fill_periodic(desc, TR, exch, mask=(1 << dim) - 2)
    assert np.array_equal(exch, want6)


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("dim,N", [(2, (24, 17)), (3, (13, 10, 12))])
def test_emulated_fused_navier_stokes_stage(dim, N, math, oracle_lib):
    """The flux-free route of a Navier-Stokes stage: fused convective stage on the six-ghost arrays, then the diffusive
    divergence accumulated on top.  Against the oracle's materialised composition (NavierStokes.cpp:2085-2092): the terms
    are the same, their association is not -- within the fast-build criterion, and ~1e-16 with reference-order kernels."""
    from common import assert_fast_parity

    desc, U = state(dim, N)
    dt = 2.0e-4
    Q6, Q4 = pb.pad_periodic(U, 6), pb.pad_periodic(U, 4)
    Fc, S = oracle_lib.compute_flux_and_source(desc, Q4, dt)
    Fd = orc.compute_diffusive_flux(desc, TR, Q6, dt)
    inner = (slice(None),) + (slice(6, -6),) * dim
    for alpha, beta in (([1.0], [1.0]), ([0.75, 0.25], [0.0, 0.25])):
        m = len(alpha)
        states = [Q6] * m if m == 1 else [pb.pad_periodic(U * 1.01, 6), Q6]
        none = [None] * (m - 1)
        want = orc.advance_stage_ns(desc, 6, alpha, beta, states, none + [Fc], none + [Fd], none + [S])[inner]
        got = emu_host.fused_stage(desc, alpha, beta, states, dt, math=math, ghosts=6)
        emu_host.diff_accumulate(desc, TR, 6, beta[-1], emu_host.diffusive_flux(desc, TR, Q6, dt), got)
        if math == 0:
            assert np.abs(got[inner] - want).max() <= 1.0e-14 * np.abs(want).max()
        else:
            assert_fast_parity(got[inner], want, "fused NS stage")


@pytest.mark.parametrize("dim,N,g", [(2, (24, 17), 6), (3, (13, 10, 12), 6), (3, (5, 1, 2), 4), (2, (1, 1), 0)])
def test_emulated_flux_free_divergence_is_the_two_call_route_bit_for_bit(dim, N, g):
    """hb2_diffusive_divergence_accumulate_dev (node fluxes of all directions kept, both faces of a cell reconstructed where
    they are differenced) against compute_diffusive_flux + accumulate: the same operations in the same order."""
    desc, U = state(dim, N)
    Q6 = pb.pad_periodic(U, 6)
    rng = np.random.default_rng(2)
    base = rng.standard_normal((desc.neq,) + tuple(n + 2 * g for n in reversed(N)))
    dt, beta = 3.0e-4, 2.0 / 3.0
    two = emu_host.diff_accumulate(desc, TR, g, beta, emu_host.diffusive_flux(desc, TR, Q6, dt), base.copy())
    one = emu_host.diff_divergence_accumulate(desc, TR, Q6, dt, g, beta, base.copy())
    assert np.array_equal(one, two)
    assert np.array_equal(one[0], base[0])                                  # no diffusive mass flux
    assert min(N) < 3 or not np.array_equal(one[1:], base[1:])


@pytest.mark.parametrize("dim,N", [(2, (24, 17)), (3, (13, 10, 12))])
def test_emulated_diffusive_spectral_radius_and_stable_dt(dim, N):
    """NavierStokes::computeSpectralRadiusesAndStableDtOnPatch: the diffusive spectral radius (emulated kernel arithmetic)
    equals the oracle's, and combined with the acoustic part (the Euler path's hb2_max_wave_speed_dev quantity, oracle:
    orc_spectral_radii_and_dt) it gives the oracle's stable dt."""
    desc, U = state(dim, N)
    Q6 = pb.pad_periodic(U, 6)
    c_p_eos = desc.gamma[0] / (desc.gamma[0] - 1.0) * 1.0
    for tr in (TR, orc.Transport(mu=1.0e-3, mu_v=0.4, c_p=3.5, c_v=2.5, Pr=0.05)):
        radii, dt, sr_diff = orc.ns_spectral_radii_and_dt(desc, tr, c_p_eos, Q6)
        assert emu_host.diff_max_spectral_radius(desc, tr, c_p_eos, Q6) == sr_diff > 0.0
        Q4 = pb.pad_periodic(U, 4)
        radii_euler, dt_euler = orc.spectral_radii_and_dt(desc, Q4)
        assert radii_euler.tolist() == radii
        want = 1.0 / (max(sr_diff, 1.0 / dt_euler) + 1.0e-15)
        assert abs(dt - want) <= 1.0e-15 * dt
    # viscosity-limited on a fine mesh: the diffusive radius takes over
    fine = orc.PatchDesc(dim=dim, n=N, gamma=desc.gamma, dx=tuple(1.0e-4 for _ in range(dim)))
    _, dt_f, sr_f = orc.ns_spectral_radii_and_dt(fine, TR, c_p_eos, Q6)
    assert abs(dt_f * (sr_f + 1.0e-15) - 1.0) < 1.0e-14


@pytest.mark.parametrize("N,g", [((13, 10, 12), 6), ((20, 7, 9), 4), ((4, 3, 5), 6)])
def test_emulated_fast_arithmetic_of_the_flux_free_route(N, g):
    """HB2_MATH_FAST of the diffusive plan (one reciprocal per cell, pre-multiplied coefficients, FMAs, energy flux from the
    momentum fluxes, analytically differenced faces) against the reference-order route: the UPDATE beta (-div F_d) agrees to
    1e-12 of its own magnitude (plus the rounding of the sum U + update itself), on a state with a Mach-3 slab."""
    desc, U = state(3, N)
    Q6 = pb.pad_periodic(U, orc.GD)
    rng = np.random.default_rng(4)
    base = rng.standard_normal((desc.neq,) + tuple(n + 2 * g for n in reversed(N)))
    dt, beta = 3.0e-4, 2.0 / 3.0
    exact = emu_host.diff_divergence_accumulate(desc, TR, Q6, dt, g, beta, base.copy())
    fast = emu_host.diff_divergence_accumulate_fast(desc, TR, Q6, dt, g, beta, base.copy())
    upd = exact - base
    assert np.abs(upd[1:]).max() > 0.0
    assert np.array_equal(fast[0], base[0])                       # no diffusive mass flux
    for e in range(1, desc.neq):
        assert np.abs(fast[e] - exact[e]).max() <= 1.0e-12 * np.abs(upd[e]).max() + 2.0 * np.finfo(float).eps * np.abs(base[e]).max(), e
    inner = (slice(None),) + (slice(g, -g),) * 3
    outside = fast.copy()
    outside[inner] = base[inner]
    assert np.array_equal(outside, base)


@pytest.mark.parametrize("dim,N", [(2, (24, 17)), (3, (13, 10, 12)), (2, (1, 1)), (3, (7, 1, 3)), (3, (4, 6, 5))])
def test_emulated_midpoint_reconstructor_is_bit_identical(dim, N):
    """DiffusiveFluxReconstructorMidpointSixthOrder: one thread forms the midpoint flux of all equations (the reference and the
    oracle stage every intermediate in its own array) -- same values bit for bit, on a state with a Mach-3 slab; ghost cells
    beyond five are never read; the continuity flux is +0.0."""
    desc, U = state(dim, N)
    Q = pb.pad_periodic(U, orc.GD)
    dt = 1.0e-3
    Fo = orc.compute_diffusive_flux_midpoint(desc, TR, Q, dt)
    Fe = emu_host.diffusive_flux_midpoint(desc, TR, Q, dt)
    for a in range(dim):
        assert np.array_equal(Fe[a], Fo[a]), f"dir {a}"
        assert not np.signbit(Fe[a][0]).any() and not Fe[a][0].any()
    Qp = Q.copy()
    outer = np.ones(Q.shape[1:], dtype=bool)
    outer[(slice(1, -1),) * dim] = False
    Qp[:, outer] = np.nan
    F2 = emu_host.diffusive_flux_midpoint(desc, TR, Qp, dt)
    assert all(np.array_equal(F2[a], Fe[a]) for a in range(dim))
    Fn = emu_host.diffusive_flux(desc, TR, Q, dt)
    if min(N) > 1:
        assert any(not np.array_equal(Fn[a], Fe[a]) for a in range(dim))      # a different discretisation of the same flux
