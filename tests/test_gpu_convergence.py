"""The reference's acceptance tests on the GPU path (configs 1-3 of BASELINE.json): L2 convergence rate > 4.8 at the
finest grid pair for WCNS5_JS_HLLC_HLL (> 4.8 WCNS5_Z, > 5.8 WCNS6_LD) with the protocol of
tests/{2D,3D}_convergence_test_*/convergence_test.py (N = 8..64, dt = 0.001*(2/8)/2^L, num_steps_base*2^L steps), and
the L1 errors of the GPU run against the oracle's: identical in the exact build (the states are bit-identical), within
1 % (L1) in the fast build.  Plus size-independent properties on a large box: discrete conservation and
translation equivariance (a periodic shift of the initial state shifts the result, bit for bit)."""
import numpy as np
import pytest

from hamers_b200 import problems as pb

pytestmark = pytest.mark.gpu


def _gpu_errors(dim, model, levels, steps_base, math, scheme=0):
    import torch
    from hamers_b200.level import UniformLevel

    out = []
    for L in range(levels):
        N = 8 * 2 ** L
        U, dx, gam = (pb.convergence_single_species(dim, N) if model == 0 else pb.convergence_five_eqn(dim, N))
        lvl = UniformLevel(dim, (N,) * dim, flow_model=model, species_gamma=gam, math=math)
        if scheme:
            from hamers_b200 import abi

            lvl.plan.close()
            lvl.plan = abi.Plan(dim, (N,) * dim, flow_model=model, species_gamma=gam, dx=lvl.dx, math=math, scheme=scheme).use_torch_stream()
        lvl.set_interior(U)
        dt = 0.001 * (2.0 / 8) / 2 ** L
        nsteps = steps_base * 2 ** L
        lvl.advance(dt, nsteps)
        torch.cuda.synchronize()
        Ug = lvl.interior().cpu().numpy()
        lvl.close()
        t = dt * nsteps
        if model == 0:
            out.append(pb.error_norms(Ug[0], pb.exact_density_single_species(dim, N, t), dx))
        else:
            out.append(pb.error_norms(Ug[-2], pb.exact_Z1_five_eqn(dim, N, t), dx))
    return out


def _oracle_errors(orc, dim, model, levels, steps_base, scheme=0):
    out = []
    for L in range(levels):
        N = 8 * 2 ** L
        U, dx, gam = (pb.convergence_single_species(dim, N) if model == 0 else pb.convergence_five_eqn(dim, N))
        lvl = orc.PatchDesc(dim=dim, n=(N,) * dim, model=model, ns=len(gam), gamma=gam, dx=dx, scheme=scheme)
        dt = 0.001 * (2.0 / 8) / 2 ** L
        nsteps = steps_base * 2 ** L
        orc.level_advance(lvl, (8,) * dim, U, dt, nsteps, nthreads=0)
        t = dt * nsteps
        if model == 0:
            out.append(pb.error_norms(U[0], pb.exact_density_single_species(dim, N, t), dx))
        else:
            out.append(pb.error_norms(U[-2], pb.exact_Z1_five_eqn(dim, N, t), dx))
    return out


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("dim,model,steps_base", [(2, 0, 8), (3, 0, 1), (3, 1, 1), (2, 1, 8)])
def test_convergence_order_and_l1_errors(dim, model, steps_base, math, oracle_lib, product_lib):
    levels = 4
    eg = _gpu_errors(dim, model, levels, steps_base, math)
    eo = _oracle_errors(oracle_lib, dim, model, levels, steps_base)
    rate = np.log2(eg[-2][1] / eg[-1][1])
    assert rate > 4.8, (eg, rate)
    for (l1g, l2g, lig), (l1o, l2o, lio) in zip(eg, eo):
        if math == 0:
            assert (l1g, l2g, lig) == (l1o, l2o, lio)
        else:
            # Per call the fast build is within 1e-12 relative of the oracle (north_star).  Over the up to 192 stages of
            # these runs the difference grows where the reference's own formulas are ill-conditioned: at smooth extrema
            # the smoothness indicators are at the level of epsilon = 1e-15 and the nonlinear weights react to
            # rounding (tests/test_oracle_conditioning.py); observed: L1 to 1e-4 relative, L2 to 6e-3, max norm to 10 %
            # of errors that are themselves 1e-8 of the solution.
            assert abs(l1g - l1o) <= 0.01 * l1o and abs(l2g - l2o) <= 0.03 * l2o and abs(lig - lio) <= 0.3 * lio


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("scheme,expected", [(1, 4.8), (2, 5.8)])
def test_convergence_order_of_the_other_interpolators(scheme, expected, math, oracle_lib, product_lib):
    eg = _gpu_errors(2, 0, 4, 8, math, scheme=scheme)
    eo = _oracle_errors(oracle_lib, 2, 0, 4, 8, scheme=scheme)
    assert np.log2(eg[-2][1] / eg[-1][1]) > expected
    if math == 0:
        assert eg == eo
    else:
        for (l1g, l2g, _), (l1o, l2o, _) in zip(eg, eo):
            assert abs(l1g - l1o) <= 0.01 * l1o and abs(l2g - l2o) <= 0.03 * l2o


@pytest.mark.parametrize("push", [False, True])
def test_large_box_conservation_and_translation_equivariance(push, product_lib):
    """192^3 (pencils much longer than the rings, several marching segments, 20+ waves of blocks): totals of the
    conserved variables are preserved to round-off, and shifting the periodic initial state shifts the result."""
    import torch
    from hamers_b200.level import UniformLevel

    N = (192, 160, 176)
    rng = np.random.default_rng(11)
    gam = (1.4,)
    z, y, x = np.meshgrid(*[np.arange(n) / n for n in reversed(N)], indexing="ij")
    rho = 1.0 + 0.3 * np.sin(2 * np.pi * (x + 2 * y)) * np.cos(2 * np.pi * z) + 0.02 * rng.standard_normal(x.shape)
    vel = [0.5 * np.sin(2 * np.pi * (y + z)), -0.7 * np.cos(2 * np.pi * x), 0.3 + 0.2 * np.sin(2 * np.pi * (x - z))]
    p = 1.0 + 0.2 * np.cos(2 * np.pi * (x + y + z))
    # a compression front so that the sensor switches HLLC-HLL on somewhere
    vel[0] = vel[0] - 1.5 * np.tanh(40.0 * (x - 0.5))
    E = p / (gam[0] - 1.0) + 0.5 * rho * (vel[0] ** 2 + vel[1] ** 2 + vel[2] ** 2)
    U = np.stack([rho, rho * vel[0], rho * vel[1], rho * vel[2], E])
    dt = 2.0e-4
    shift = (101, 5, 37)      # cells in (z, y, x)

    def run(U0):
        lvl = UniformLevel(3, N, species_gamma=gam, math=1, push=push)
        lvl.set_interior(U0)
        tot0 = lvl.interior().sum(dim=(1, 2, 3)).cpu().numpy()
        lvl.advance(dt, 2)
        torch.cuda.synchronize()
        out = lvl.interior().cpu().numpy()
        tot1 = lvl.interior().sum(dim=(1, 2, 3)).cpu().numpy()
        lvl.close()
        return out, tot0, tot1

    A, t0, t1 = run(U)
    scale = np.abs(U).sum(axis=(1, 2, 3))
    assert np.all(np.abs(t1 - t0) <= 1e-12 * scale), (t0, t1)
    assert np.isfinite(A).all()
    B, _, _ = run(np.ascontiguousarray(np.roll(U, shift, axis=(1, 2, 3))))
    assert np.array_equal(B, np.roll(A, shift, axis=(1, 2, 3)))


def test_full_size_512_conservation_and_translation_equivariance(product_lib):
    """BASELINE.json's full size (512^3, the bench workload's box), everything on the device: totals of the conserved
    variables are preserved to round-off over a step, and a periodic shift of the initial state shifts the result bit
    for bit (exercises 17-chunk x pencils, 65-chunk y/z pencils, multi-segment sensor marches, 30+ waves of blocks)."""
    import torch
    from hamers_b200.level import UniformLevel

    if torch.cuda.get_device_properties(0).total_memory < 80e9:
        pytest.skip("needs ~60 GB of device memory")
    N = 512
    g = torch.Generator(device="cuda").manual_seed(5)
    ax = (torch.arange(N, device="cuda", dtype=torch.float64) + 0.5) / N
    x, y, z = ax[None, None, :], ax[None, :, None], ax[:, None, None]
    two_pi = 2.0 * np.pi
    rho = 1.0 + 0.3 * torch.sin(two_pi * (x + 2 * y)) * torch.cos(two_pi * z) + 0.01 * torch.randn((N, N, N), generator=g, device="cuda", dtype=torch.float64)
    u = 0.5 * torch.sin(two_pi * (y + z)) - 1.5 * torch.tanh(40.0 * (x - 0.5)) + 0 * z
    v = -0.7 * torch.cos(two_pi * x) + 0 * y + 0 * z
    w = 0.3 + 0.2 * torch.sin(two_pi * (x - z)) + 0 * y
    p = 1.0 + 0.2 * torch.cos(two_pi * (x + y + z))
    E = p / 0.4 + 0.5 * rho * (u * u + v * v + w * w)
    U = torch.stack([rho, rho * u, rho * v, rho * w, E]).contiguous()
    del rho, u, v, w, p, E
    shift = (101, 5, 37)
    dt = 0.25 * 0.001 * 2.0 / N

    def run(U0):
        lvl = UniformLevel(3, (N, N, N), species_gamma=(1.4,), math=1)
        lvl.interior().copy_(U0)
        lvl.advance(dt, 1)
        torch.cuda.synchronize()
        out = lvl.interior().clone()
        lvl.close()
        return out

    A = run(U)
    assert bool(torch.isfinite(A).all())
    t0, t1 = U.sum(dim=(1, 2, 3)), A.sum(dim=(1, 2, 3))
    scale = U.abs().sum(dim=(1, 2, 3))
    assert bool(((t1 - t0).abs() <= 1e-12 * scale).all()), (t0, t1)
    B = run(torch.roll(U, shift, dims=(1, 2, 3)).contiguous())
    assert torch.equal(B, torch.roll(A, shift, dims=(1, 2, 3)))
