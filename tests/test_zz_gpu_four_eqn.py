"""SURVEY row f3 on the GPU: the FOUR_EQN_CONSERVATIVE flow model through the C ABI against the oracle -- bit-identical
(the model runs the reference-order kernels whatever `math` says), 2-D / 3-D, three interpolators; spectral radii; the
reference's acceptance protocol (L2 rate > 4.8 at the finest grid pair) with two species; the RMI state of config 4."""
import numpy as np
import pytest

from hamers_b200 import problems as pb

pytestmark = pytest.mark.gpu
FC = 2


def _plan(desc, math=0):
    from hamers_b200 import abi

    return abi.Plan(desc.dim, desc.n, flow_model=abi.FOUR_EQN_CONSERVATIVE, species_gamma=desc.gamma, species_R=desc.R,
                    dx=desc.dx, math=math, scheme=desc.scheme).use_torch_stream()


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("scheme", [0, 1, 2])
@pytest.mark.parametrize("dim,N", [(2, (24, 17)), (3, (13, 10, 12))])
def test_four_eqn_flux_stage_and_wave_speed(dim, N, scheme, math, oracle_lib, product_lib):
    import torch

    U, dx, gam, R = pb.random_state_four_eqn(dim, N)
    desc = oracle_lib.PatchDesc(dim=dim, n=N, model=FC, ns=2, gamma=gam, R=R, dx=dx, scheme=scheme)
    Q = pb.pad_periodic(U)
    dt = 1.0e-3
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt)
    Uo = oracle_lib.advance_stage(desc, [1.0], [1.0], [Q], [Fo], [So])
    plan = _plan(desc, math)
    assert plan.neq == dim + 3 and plan.ncomp == plan.neq
    Qd = torch.from_numpy(Q).cuda()
    Fd = [torch.full((desc.neq,) + desc.side_shape(a), float("nan"), dtype=torch.float64, device="cuda") for a in range(dim)]
    Sd = torch.zeros((desc.neq,) + desc.cell_shape, dtype=torch.float64, device="cuda")
    plan.compute_flux_and_source(Qd, dt, Fd, Sd)
    out = torch.zeros_like(Qd)
    plan.fused_stage([1.0], [1.0], [Qd], dt, out)
    # materialised route of the same stage (what the AMR flux synchronisation uses)
    out2 = torch.zeros_like(Qd)
    plan.advance_stage([1.0], [1.0], [Qd], [Fd], [Sd], out2)
    sr = torch.zeros(4, dtype=torch.float64, device="cuda")
    plan.max_wave_speed(Qd, sr)
    torch.cuda.synchronize()
    for a in range(dim):
        assert np.array_equal(Fd[a].cpu().numpy(), Fo[a]), f"dir {a}"
    inner = (slice(None),) + (slice(4, -4),) * dim
    assert np.array_equal(out.cpu().numpy()[inner], Uo[inner])
    assert np.array_equal(out2.cpu().numpy()[inner], Uo[inner])
    want, dt_o = oracle_lib.spectral_radii_and_dt(desc, Q, include_ghosts=False)
    assert np.array_equal(sr.cpu().numpy()[:dim], want) and 1.0 / float(sr[3]) == dt_o
    plan.close()


def test_four_eqn_convergence_order_on_the_gpu(oracle_lib, product_lib):
    import torch
    from hamers_b200 import abi
    from hamers_b200.level import UniformLevel

    errs, errs_o = [], []
    for L in range(4):
        N = 8 * 2 ** L
        U, dx, gam, R = pb.convergence_four_eqn(2, N)
        lvl = UniformLevel(2, (N, N), flow_model=abi.FOUR_EQN_CONSERVATIVE, species_gamma=gam, species_R=R, math=abi.MATH_EXACT)
        lvl.set_interior(U)
        dt = 0.001 * (2.0 / 8) / 2 ** L
        nsteps = 8 * 2 ** L
        lvl.advance(dt, nsteps)
        torch.cuda.synchronize()
        Ug = lvl.interior().cpu().numpy()
        lvl.close()
        errs.append(pb.error_norms(Ug[0], pb.exact_rhoY1_four_eqn(2, N, dt * nsteps), dx))
        d = oracle_lib.PatchDesc(dim=2, n=(N, N), model=FC, ns=2, gamma=gam, R=R, dx=dx)
        oracle_lib.level_advance(d, (8, 8), U, dt, nsteps, nthreads=0)
        errs_o.append(pb.error_norms(U[0], pb.exact_rhoY1_four_eqn(2, N, dt * nsteps), dx))
    assert np.log2(errs[-2][1] / errs[-1][1]) > 4.8, errs
    assert errs == errs_o            # the states are bit-identical, so are the norms


def test_richtmyer_meshkov_state_of_config_4(oracle_lib, product_lib):
    """One stage on the shocked SF6 / air state of the shipped deck (WCNS6_LD_HLLC_HLL, species constants of the deck), on a
    periodic copy of the box: bit-identical to the oracle, sensor and fallback faces included."""
    import torch

    N = (128, 16)
    U, dx, gam, R = pb.richtmyer_meshkov_2d(N, x_up=(0.004, 0.0005))
    desc = oracle_lib.PatchDesc(dim=2, n=N, model=FC, ns=2, gamma=gam, R=R, dx=dx, scheme=2)
    Q = pb.pad_periodic(U)
    dt = 1.0e-9
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt)
    plan = _plan(desc)
    Qd = torch.from_numpy(Q).cuda()
    Fd = [torch.empty((desc.neq,) + desc.side_shape(a), dtype=torch.float64, device="cuda") for a in range(2)]
    Sd = torch.zeros((desc.neq,) + desc.cell_shape, dtype=torch.float64, device="cuda")
    plan.compute_flux_and_source(Qd, dt, Fd, Sd)
    torch.cuda.synchronize()
    for a in range(2):
        assert np.isfinite(Fo[a]).all() and np.array_equal(Fd[a].cpu().numpy(), Fo[a])
    plan.close()
