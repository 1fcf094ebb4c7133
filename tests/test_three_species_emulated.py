"""Three species (the reference's flow models are generic in d_num_species; FlowModelFiveEqnAllaire.cpp:29,
FlowModelFourEqnConservative.cpp:29): the five-eqn and four-eqn instantiations of the kernels' thread functions with
NS = 3 -- the five-eqn ones in COMPACT 128-thread blocks (half the pencils: the rings of a 256-thread block would need
273 KB of shared memory) -- against the oracle, bit for bit, on the CPU (tests/host_emu compiles the same
`__host__ __device__` code with g++).  Oracle status at ns = 3: its HLLC / HLLC-HLL kernels are pinned against the
reference's own compiled ones at three species (test_riemann_kernels_at_three_species_match_the_reference); the mixture
chains are the loops pinned at two species run over one more species."""
import numpy as np
import pytest

import emu_host
from hamers_b200 import problems as pb
from oracle import oracle as orc

FE, FC = orc.FIVE_EQN_ALLAIRE, orc.FOUR_EQN_CONSERVATIVE


def _case(model, dim, N, scheme=0, seed=20261017):
    if model == FE:
        U, dx, gam = pb.random_state_three_species(dim, N, FE, seed=seed)
        return orc.PatchDesc(dim=dim, n=N, model=FE, ns=3, gamma=gam, dx=dx, scheme=scheme), U
    U, dx, gam, R = pb.random_state_three_species(dim, N, FC, seed=seed)
    return orc.PatchDesc(dim=dim, n=N, model=FC, ns=3, gamma=gam, R=R, dx=dx, scheme=scheme), U


def _stage_states(desc, U, m):
    """m - 1 older states + the flux state; the five-eqn model stores the derived last volume fraction"""
    out = []
    for k in range(m - 1):
        V = U * (1.0 + 0.01 * k)
        if desc.model == FE:
            V[-desc.ns:] = U[-desc.ns:]
        out.append(pb.pad_periodic(V))
    return out + [pb.pad_periodic(U)]


@pytest.mark.parametrize("scheme", [0, 1, 2])
@pytest.mark.parametrize("model,dim,N", [(FE, 2, (24, 17)), (FE, 3, (13, 10, 12)), (FC, 2, (24, 17)), (FC, 3, (13, 10, 12)),
                                         (FE, 3, (35, 4, 5))])
def test_emulated_three_species_flux_and_stage_match_oracle(model, dim, N, scheme, oracle_lib):
    desc, U = _case(model, dim, N, scheme)
    Q = pb.pad_periodic(U)
    dt = 1.0e-3
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt)
    assert all(np.isfinite(f).all() for f in Fo)
    Fe, Se = emu_host.flux_and_source(desc, Q, dt, math=0)
    for a in range(dim):
        assert np.array_equal(Fe[a], Fo[a]), f"dir {a}"
    assert np.array_equal(Se, So)
    assert So.any() == (model == FE)                  # only the five-eqn model has advective equations
    inner = (slice(None),) + (slice(4, -4),) * dim
    for alpha, beta in (([1.0], [1.0]), ([0.75, 0.25], [0.0, 0.25]), ([1.0 / 3.0, 0.0, 2.0 / 3.0], [0.0, 0.0, 2.0 / 3.0])):
        m = len(alpha)
        states = _stage_states(desc, U, m)
        none = [None] * (m - 1)
        Uo = oracle_lib.advance_stage(desc, alpha, beta, states, none + [Fo], none + [So])
        Ue = emu_host.fused_stage(desc, alpha, beta, states, dt, math=0)
        assert np.array_equal(Ue[inner], Uo[inner])


@pytest.mark.parametrize("model", [FE, FC])
@pytest.mark.parametrize("N", [(4, 4, 4), (5, 7, 6), (17, 4, 5)])
def test_emulated_three_species_tiny_patches_and_push(model, N, oracle_lib):
    desc, U = _case(model, 3, N, seed=3)
    Q = pb.pad_periodic(U)
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, 1.0e-3)
    Uo = oracle_lib.advance_stage(desc, [1.0], [1.0], [Q], [Fo], [So])
    Ue = emu_host.fused_stage(desc, [1.0], [1.0], [Q], 1.0e-3, math=0, push=True)
    inner = (slice(None),) + (slice(4, -4),) * 3
    assert np.array_equal(Ue[inner], Uo[inner])
    assert np.array_equal(Ue, pb.pad_periodic(np.ascontiguousarray(Ue[inner])))
