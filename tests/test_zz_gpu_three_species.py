"""Three species on the GPU (the reference's flow models are generic in d_num_species): FIVE_EQN_ALLAIRE and
FOUR_EQN_CONSERVATIVE with num_species = 3 through the C ABI against the oracle -- bit-identical (three species run the
reference-order kernels whatever `math` says; the five-eqn ones in compact 128-thread blocks), 2-D / 3-D, three interpolators,
materialised and fused stage (all three SSP-RK3 rows), spectral radii, multi-segment pencils."""
import numpy as np
import pytest

from hamers_b200 import problems as pb

pytestmark = pytest.mark.gpu
FE, FC = 1, 2


def _case(oracle_lib, model, dim, N, scheme, seed=20261017):
    if model == FE:
        U, dx, gam = pb.random_state_three_species(dim, N, FE, seed=seed)
        return oracle_lib.PatchDesc(dim=dim, n=N, model=FE, ns=3, gamma=gam, dx=dx, scheme=scheme), U
    U, dx, gam, R = pb.random_state_three_species(dim, N, FC, seed=seed)
    return oracle_lib.PatchDesc(dim=dim, n=N, model=FC, ns=3, gamma=gam, R=R, dx=dx, scheme=scheme), U


def _plan(desc, math=0):
    from hamers_b200 import abi

    kw = dict(species_R=desc.R) if desc.model == FC else {}
    return abi.Plan(desc.dim, desc.n, flow_model=desc.model, species_gamma=desc.gamma, dx=desc.dx, math=math, scheme=desc.scheme,
                    **kw).use_torch_stream()


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("scheme", [0, 1, 2])
@pytest.mark.parametrize("model,dim,N", [(FE, 2, (24, 17)), (FE, 3, (13, 10, 12)), (FC, 2, (24, 17)), (FC, 3, (13, 10, 12)),
                                         (FE, 3, (70, 5, 4)), (FE, 2, (40, 150))])
def test_three_species_flux_stage_and_wave_speed(model, dim, N, scheme, math, oracle_lib, product_lib):
    import torch

    desc, U = _case(oracle_lib, model, dim, N, scheme)
    Q = pb.pad_periodic(U)
    dt = 1.0e-3
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt)
    plan = _plan(desc, math)
    assert plan.neq == desc.neq and plan.ncomp == Q.shape[0]
    Qd = torch.from_numpy(Q).cuda()
    Fd = [torch.full((desc.neq,) + desc.side_shape(a), float("nan"), dtype=torch.float64, device="cuda") for a in range(dim)]
    Sd = torch.zeros((desc.neq,) + desc.cell_shape, dtype=torch.float64, device="cuda")
    plan.compute_flux_and_source(Qd, dt, Fd, Sd)
    torch.cuda.synchronize()
    for a in range(dim):
        assert np.array_equal(Fd[a].cpu().numpy(), Fo[a]), f"dir {a}"
    assert np.array_equal(Sd.cpu().numpy(), So)
    inner = (slice(None),) + (slice(4, -4),) * dim
    for alpha, beta in (([1.0], [1.0]), ([0.75, 0.25], [0.0, 0.25]), ([1.0 / 3.0, 0.0, 2.0 / 3.0], [0.0, 0.0, 2.0 / 3.0])):
        m = len(alpha)
        older = []
        for k in range(m - 1):
            V = U * (1.0 + 0.01 * k)
            if model == FE:
                V[-3:] = U[-3:]
            older.append(pb.pad_periodic(V))
        states = older + [Q]
        none = [None] * (m - 1)
        Uo = oracle_lib.advance_stage(desc, alpha, beta, states, none + [Fo], none + [So])
        sd = [torch.from_numpy(s).cuda() for s in states]
        out = torch.zeros_like(Qd)
        plan.fused_stage(alpha, beta, sd, dt, out)
        out2 = torch.zeros_like(Qd)
        plan.advance_stage(alpha, beta, sd, none + [Fd], none + [Sd], out2)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy()[inner], Uo[inner]), f"fused, {m} states"
        assert np.array_equal(out2.cpu().numpy()[inner], Uo[inner]), f"materialised, {m} states"
    sr = torch.zeros(4, dtype=torch.float64, device="cuda")
    plan.max_wave_speed(Qd, sr)
    torch.cuda.synchronize()
    want, dt_o = oracle_lib.spectral_radii_and_dt(desc, Q, include_ghosts=False)
    assert np.array_equal(sr.cpu().numpy()[:dim], want) and 1.0 / float(sr[3]) == dt_o
    plan.close()


@pytest.mark.parametrize("model", [FE, FC])
def test_three_species_level_steps_match_the_oracle(model, oracle_lib, product_lib):
    """Two SSP-RK3 steps of the GPU-resident level (ghost fills included) against the oracle's level loop."""
    import torch
    from hamers_b200.level import UniformLevel

    N = (20, 12, 9)
    desc, U = _case(oracle_lib, model, 3, N, 0, seed=9)
    kw = dict(species_R=desc.R) if model == FC else {}
    lvl = UniformLevel(3, N, flow_model=model, species_gamma=desc.gamma, math=0, **kw)
    lvl.set_interior(U)
    dt = 2.0e-4
    lvl.advance(dt, 2)
    torch.cuda.synchronize()
    got = lvl.interior().cpu().numpy()
    want = oracle_lib.level_advance(desc, N, np.ascontiguousarray(U.copy()), dt, 2)
    assert np.array_equal(got, want)
    lvl.close()
