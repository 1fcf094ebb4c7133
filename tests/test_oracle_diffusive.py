"""SURVEY.md row f4, CPU side: the oracle of the node-based sixth-order diffusive flux (oracle/oracle_diffusive.c).

* physics: on a smooth periodic field the divergence of the reconstructed side flux converges at sixth order to the
  divergence of the exact viscous flux (velocity / temperature derivatives taken spectrally);
* structure: conservation, the zero flux of a uniform state, equivariance under the swap of two axes;
* the point formulas and kernels against the reference's own code: tests/test_oracle_pinned.py (golden fixture)."""
import numpy as np
import pytest

from hamers_b200 import problems as pb
from oracle import oracle as orc

GAMMA = 1.4
TR = orc.Transport(mu=0.05, mu_v=0.02, c_p=3.5, c_v=2.5, Pr=0.72)


def smooth_state(dim, N):
    """Periodic, smooth, all quantities varying in all directions; box [0, 1)^dim."""
    ax = [(np.arange(n) + 0.5) / n for n in N]
    X = np.meshgrid(*reversed(ax), indexing="ij")[::-1]          # X[a] has shape (z, y, x)
    two_pi = 2.0 * np.pi
    ph = sum(X[a] for a in range(dim))
    rho = 1.0 + 0.1 * np.sin(two_pi * ph)
    vel = [0.5 * np.cos(two_pi * (X[a] + 0.3 * a)) * np.sin(two_pi * X[(a + 1) % dim]) + 0.1 * a for a in range(dim)]
    p = 1.0 + 0.3 * np.cos(two_pi * X[0]) * np.cos(two_pi * X[dim - 1])
    E = p / (GAMMA - 1.0) + 0.5 * rho * sum(v * v for v in vel)
    return np.stack([rho] + [rho * v for v in vel] + [E]), rho, vel, p


def spectral_derivative(f, axis, length=1.0):
    n = f.shape[axis]
    k = 2.0j * np.pi * np.fft.fftfreq(n, d=length / n)
    shape = [1] * f.ndim
    shape[axis] = n
    return np.real(np.fft.ifft(np.fft.fft(f, axis=axis) * k.reshape(shape), axis=axis))


def exact_viscous_divergence(dim, rho, vel, p):
    """div F_d of the exact flux at the cell centres, F_d = -(0, tau . e_d, tau . u + kappa grad T)."""
    mu, mu_v = TR.mu, TR.mu_v
    kappa = TR.c_p * mu / TR.Pr
    T = p / ((GAMMA - 1.0) * TR.c_v * rho)
    np_axis = lambda a: dim - 1 - a                                 # noqa: E731  (x is the fastest numpy axis)
    grad = [[spectral_derivative(vel[i], np_axis(j)) for j in range(dim)] for i in range(dim)]
    div_u = sum(grad[i][i] for i in range(dim))
    tau = [[mu * (grad[i][j] + grad[j][i]) + ((mu_v - 2.0 / 3.0 * mu) * div_u if i == j else 0.0) for j in range(dim)]
           for i in range(dim)]
    out = [np.zeros_like(rho)]
    for i in range(dim):
        out.append(-sum(spectral_derivative(tau[i][j], np_axis(j)) for j in range(dim)))
    q = [sum(tau[i][j] * vel[i] for i in range(dim)) + kappa * spectral_derivative(T, np_axis(j)) for j in range(dim)]
    out.append(-sum(spectral_derivative(q[j], np_axis(j)) for j in range(dim)))
    return np.stack(out)


def flux_divergence(desc, F, dt):
    dim = desc.dim
    div = np.zeros((desc.neq,) + desc.cell_shape)
    for a in range(dim):
        ax = dim - a                                              # numpy axis of direction a in (neq, z, y, x)
        hi = [slice(None)] * (dim + 1)
        lo = [slice(None)] * (dim + 1)
        hi[ax], lo[ax] = slice(1, None), slice(0, -1)
        div += (F[a][tuple(hi)] - F[a][tuple(lo)]) / desc.dx[a] / dt
    return div


def run(dim, N, dt=1.0e-3):
    U, rho, vel, p = smooth_state(dim, N)
    desc = orc.PatchDesc(dim=dim, n=N, gamma=(GAMMA,), dx=tuple(1.0 / n for n in N))
    F = orc.compute_diffusive_flux(desc, TR, pb.pad_periodic(U, orc.GD), dt)
    return desc, F, flux_divergence(desc, F, dt), exact_viscous_divergence(dim, rho, vel, p)


@pytest.mark.parametrize("dim", [2, 3])
def test_flux_divergence_converges_at_sixth_order(dim):
    errs = []
    for n in (16, 32):
        N = (n, n + 4, n + 2)[:dim]
        _, _, div, exact = run(dim, N)
        assert np.abs(div[0]).max() == 0.0                         # no diffusive mass flux
        errs.append(np.abs(div[1:] - exact[1:]).max())
    order = np.log2(errs[0] / errs[1])
    assert errs[1] < 1.0e-4 * np.abs(exact[1:]).max() and order > 5.3, (errs, order)


@pytest.mark.parametrize("dim", [2, 3])
def test_conservation_uniform_state_and_axis_swap(dim):
    N = (10, 12, 9)[:dim]
    desc, F, div, _ = run(dim, N)
    scale = np.abs(div).max()
    assert np.abs(div.reshape(desc.neq, -1).sum(axis=1)).max() < 1.0e-11 * scale * div[0].size      # telescoping sum
    # periodic wrap: first and last face of every line carry the same flux
    for a in range(dim):
        ax = dim - a
        first = np.take(F[a], 0, axis=ax)
        last = np.take(F[a], -1, axis=ax)
        assert np.array_equal(first, last)
    # uniform state: every derivative is exactly zero
    U0 = np.ones((desc.neq,) + desc.cell_shape) * np.array([1.3, 0.2, -0.4, 0.7, 5.0][:dim + 1] + [5.0]).reshape(
        (-1,) + (1,) * dim)
    F0 = orc.compute_diffusive_flux(desc, TR, pb.pad_periodic(U0, orc.GD), 1.0e-3)
    assert all(np.abs(f).max() == 0.0 for f in F0)
    # swapping the x and y axes of the state (and of the momentum components) swaps the fluxes: the term tables of the
    # two directions are consistent with each other
    U, *_ = smooth_state(dim, N)
    perm = [0, 2, 1] + list(range(3, desc.neq))
    Us = np.swapaxes(U, -1, -2)[perm]
    Ns = (N[1], N[0]) + tuple(N[2:])
    descs = orc.PatchDesc(dim=dim, n=Ns, gamma=(GAMMA,), dx=tuple(1.0 / n for n in Ns))
    Fs = orc.compute_diffusive_flux(descs, TR, pb.pad_periodic(np.ascontiguousarray(Us), orc.GD), 1.0e-3)
    for a, b in ((0, 1), (1, 0)) + (((2, 2),) if dim == 3 else ()):
        got = np.swapaxes(Fs[b], -1, -2)[perm]
        assert np.allclose(got, F[a], rtol=0.0, atol=1.0e-13 * np.abs(F[a]).max()), (a, b)


def test_stage_update_combines_both_fluxes_like_the_reference():
    """NavierStokes.cpp:2085-2092: -(Fc_R - Fc_L + Fd_R - Fd_L)/dx per direction, in that association."""
    rng = np.random.default_rng(3)
    N, g = (5, 4, 3), 6
    desc = orc.PatchDesc(dim=3, n=N, gamma=(GAMMA,), dx=(0.1, 0.2, 0.3))
    shape = tuple(n + 2 * g for n in reversed(N))
    U = [rng.standard_normal((5,) + shape) for _ in range(2)]
    Fc = [[rng.standard_normal((5,) + desc.side_shape(a)) for a in range(3)] for _ in range(2)]
    Fd = [[rng.standard_normal((5,) + desc.side_shape(a)) for a in range(3)] for _ in range(2)]
    S = [rng.standard_normal((5,) + desc.cell_shape) for _ in range(2)]
    alpha, beta = [0.75, 0.25], [0.0, 0.25]
    out = orc.advance_stage_ns(desc, g, alpha, beta, U, Fc, Fd, S)
    inner = (slice(None),) + (slice(g, -g),) * 3
    ref = np.zeros((5,) + desc.cell_shape)
    ref += alpha[0] * U[0][inner]
    ref += alpha[1] * U[1][inner]
    d = lambda F, ax: (np.diff(F, axis=ax))                          # noqa: E731
    m = 1
    ref += beta[m] * (-(Fc[m][0][..., 1:] - Fc[m][0][..., :-1] + Fd[m][0][..., 1:] - Fd[m][0][..., :-1]) / desc.dx[0]
                      - (Fc[m][1][..., 1:, :] - Fc[m][1][..., :-1, :] + Fd[m][1][..., 1:, :] - Fd[m][1][..., :-1, :]) / desc.dx[1]
                      - (Fc[m][2][:, 1:] - Fc[m][2][:, :-1] + Fd[m][2][:, 1:] - Fd[m][2][:, :-1]) / desc.dx[2] + S[m])
    assert np.array_equal(out[inner], ref)
    ghost = out.copy()
    ghost[inner] = 0.0
    assert np.abs(ghost).max() == 0.0


# ---- pinned against the reference's own kernels (tests/golden/make_golden_diffusive.py) -----------------------------
import os  # noqa: E402

DGOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "diffusive_kernels.npz"))


@pytest.mark.parametrize("dim", [2, 3])
def test_derivative_and_reconstruction_kernels_match_reference_golden(dim):
    """DiffusiveFluxReconstructorNodeSixthOrder::computeFirstDerivativesIn{X,Y,Z} and reconstructFlux{X,Y,Z}, compiled
    verbatim as members of a stub class and driven over the base class's index ranges: the oracle's array kernels give
    the same numbers AND write the same set of entries (NaN elsewhere)."""
    n = (5, 4, 3)[:dim]
    u = DGOLD[f"u{dim}d"]
    for d in range(dim):
        der = orc.diff_derivative_array(dim, d, u, n, 7.3)
        ref = DGOLD[f"der{dim}d{d}"]
        assert np.array_equal(np.isnan(der), np.isnan(ref))
        assert np.array_equal(der[~np.isnan(ref)], ref[~np.isnan(ref)])
        assert (~np.isnan(ref)).sum() == np.prod([x + 12 - (6 if a == d else 0) for a, x in enumerate(n)])
        rec = orc.diff_reconstruct_array(dim, d, u, n, 0.37)
        assert np.array_equal(rec, DGOLD[f"rec{dim}d{d}"])


@pytest.mark.parametrize("dim", [2, 3])
def test_temperature_conductivity_and_diffusivities_match_reference_golden(dim):
    """T (EquationOfStateIdealGas.cpp:6897), kappa (EquationOfThermalConductivityPrandtl.cpp:309) and D_00.. of
    FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:4180-4193 / 4262-4280: the reference's statements compiled verbatim."""
    for v, ref in zip(DGOLD[f"point{dim}d_in"], DGOLD[f"point{dim}d_out"]):
        tr = orc.Transport(mu=v[5], mu_v=v[7], c_p=v[4], c_v=v[1], Pr=v[6])
        T, kappa, D = orc.diff_point(dim, v[0], v[2], v[3], v[8:8 + dim], tr)
        assert [T, kappa] + D == ref.tolist()


@pytest.mark.parametrize("dim", [2, 3])
def test_max_diffusivity_spectral_radius_and_dt_match_reference_golden(dim):
    """MAX_DIFFUSIVITY (FlowModelSingleSpecies.cpp:4661-4665), the diffusive spectral radius and the stable dt of
    NavierStokes::computeSpectralRadiusesAndStableDtOnPatch (NavierStokes.cpp:884-893, 1083-1091): the reference's statements
    compiled verbatim; outputs in the golden fixture."""
    for v, ref in zip(DGOLD[f"dt{dim}d_in"], DGOLD[f"dt{dim}d_out"]):
        D, radius = orc.diff_max_diffusivity_point(dim, v[0], v[1], v[2], v[3], v[4], v[5:5 + dim])
        dt = 1.0 / (max(radius, v[8]) + 1.0e-15)
        assert [D, radius, dt] == ref.tolist()


@pytest.mark.parametrize("dim", [2, 3])
def test_term_tables_match_reference_golden(dim):
    """Which derivative carries which diffusivity in which equation, in the order of accumulation: the oracle's tables and
    the compile-time tables of the product kernels (hb2_diffusive.cuh: DiffTerms) against the reference's own
    getCellDataOfDiffusiveFluxVariablesForDerivative / getCellDataOfDiffusiveFluxDiffusivities, compiled verbatim as members
    of a stub class (oracle/build_ref.py: diffusive_term_tables; outputs in the golden fixture)."""
    import emu_host

    ref = DGOLD[f"terms{dim}d"]
    nterms = 0
    for f in range(dim):
        for d in range(dim):
            for e in range(dim + 2):
                want = [tuple(int(x) for x in t) for t in ref[f, d, e] if t[0] >= 0]
                assert orc.diff_terms(dim, f, d, e) == want, (f, d, e)
                assert emu_host.diff_terms(dim, f, d, e) == want, (f, d, e)
                nterms += len(want)
    assert nterms == (45 if dim == 3 else 18)       # 3 x (7 + 4 + 4) and 2 x (5 + 4)


def test_term_tables_are_the_stress_tensor():
    """Independent cross-check of the term tables (pinned in the test above): assembled symbolically they give F_d = -(0, tau . e_f, u . tau . e_f + kappa dT/dx_f) with the Newtonian tau, for every direction."""
    for dim in (2, 3):
        rng = np.random.default_rng(dim)
        mu, mu_v, kappa = 0.3, 0.07, 1.9
        vel = rng.standard_normal(3)
        grad = rng.standard_normal((dim + 1, dim))                 # grad[var][direction]; var = velocity comps, T
        _, _, D = orc.diff_point(dim, 1.4, 1.0, 1.0, vel[:dim], orc.Transport(mu=mu, mu_v=mu_v, c_p=kappa, c_v=1.0, Pr=mu))
        assert D[-1] == -kappa
        g = grad[:dim]
        div = np.trace(g)
        tau = mu * (g + g.T) + (mu_v - 2.0 / 3.0 * mu) * div * np.eye(dim)
        for f in range(dim):
            want = np.concatenate([[0.0], -tau[:, f], [-(vel[:dim] @ tau[:, f]) - kappa * grad[dim, f]]])
            for e in range(dim + 2):
                got = sum(D[k] * grad[var, d] for d in range(dim) for var, k in orc.diff_terms(dim, f, d, e))
                assert abs(got - want[e]) < 1.0e-13, (dim, f, e, got, want[e])


def test_taylor_green_vortex_dissipates_at_the_analytic_rate():
    """End-to-end sign and magnitude check of the Navier-Stokes composition (convective flux + diffusive flux + conservative
    stage update, SSP-RK3): at t = 0 the Taylor-Green vortex loses kinetic energy at dE_k/dt = -nu <|omega|^2> = -3 nu / 4.
    The inviscid run of the same composition carries the scheme's own dissipation and the compressible pressure work
    (M = 0.1), so the viscous contribution is the difference of the two."""
    from test_ns_level_emulated import oracle_ns_step

    N, L, M0, nu = (16, 16, 16), 2.0 * np.pi, 0.1, 0.02
    ax = [(np.arange(n) + 0.5) * L / n for n in N]
    Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
    u, v = np.sin(X) * np.cos(Y) * np.cos(Z), -np.cos(X) * np.sin(Y) * np.cos(Z)
    p = 1.0 / (1.4 * M0 * M0) + (np.cos(2 * X) + np.cos(2 * Y)) * (np.cos(2 * Z) + 2.0) / 16.0
    rho = np.ones_like(u)
    U = np.stack([rho, rho * u, rho * v, np.zeros_like(u), p / 0.4 + 0.5 * rho * (u * u + v * v)])
    desc = orc.PatchDesc(dim=3, n=N, gamma=(1.4,), dx=tuple(L / n for n in N))
    dt = 0.3 * desc.dx[0] / (1.0 + 1.0 / M0)

    def rate(mu):
        tr = orc.Transport(mu=mu, mu_v=0.0, c_p=3.5, c_v=2.5, Pr=0.71)
        W = U
        for _ in range(6):
            W = oracle_ns_step(desc, tr, W, dt)
        ke = lambda A: float((0.5 * (A[1] ** 2 + A[2] ** 2 + A[3] ** 2) / A[0]).mean())      # noqa: E731
        return (ke(W) - ke(U)) / (6 * dt)

    viscous = rate(nu) - rate(1.0e-12)
    assert abs(viscous / (-0.75 * nu) - 1.0) < 0.03, viscous
