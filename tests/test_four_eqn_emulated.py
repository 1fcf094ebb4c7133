"""SURVEY row f3, product side on the CPU: the four-eqn conservative instantiation of the kernels' thread functions
(tests/host_emu compiles the same `__host__ __device__` code with g++) against the oracle, bit for bit -- fluxes, fused
stage, the three interpolators, 2-D / 3-D, tiny patches with the fused ghost push."""
import numpy as np
import pytest

import emu_host
from hamers_b200 import problems as pb
from oracle import oracle as orc

FC = orc.FOUR_EQN_CONSERVATIVE


def _case(dim, N, scheme=0, seed=20261017):
    U, dx, gam, R = pb.random_state_four_eqn(dim, N, seed=seed)
    return orc.PatchDesc(dim=dim, n=N, model=FC, ns=2, gamma=gam, R=R, dx=dx, scheme=scheme), U


@pytest.mark.parametrize("scheme", [0, 1, 2])
@pytest.mark.parametrize("dim,N", [(2, (24, 17)), (3, (13, 10, 12))])
def test_emulated_four_eqn_flux_and_stage_match_oracle(dim, N, scheme, oracle_lib):
    desc, U = _case(dim, N, scheme)
    Q = pb.pad_periodic(U)
    dt = 1.0e-3
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt)
    assert all(np.isfinite(f).all() for f in Fo) and not So.any()          # no advective equation: the source stays zero
    Fe, Se = emu_host.flux_and_source(desc, Q, dt, math=0)
    for a in range(dim):
        assert np.array_equal(Fe[a], Fo[a]), f"dir {a}"
    assert not Se.any()
    inner = (slice(None),) + (slice(4, -4),) * dim
    for alpha, beta in (([1.0], [1.0]), ([0.75, 0.25], [0.0, 0.25]), ([1.0 / 3.0, 0.0, 2.0 / 3.0], [0.0, 0.0, 2.0 / 3.0])):
        m = len(alpha)
        states = [pb.pad_periodic(U * (1.0 + 0.01 * k)) for k in range(m - 1)] + [Q]
        none = [None] * (m - 1)
        Uo = oracle_lib.advance_stage(desc, alpha, beta, states, none + [Fo], none + [So])
        Ue = emu_host.fused_stage(desc, alpha, beta, states, dt, math=0)
        assert np.array_equal(Ue[inner], Uo[inner])


@pytest.mark.parametrize("N", [(4, 4, 4), (5, 7, 6), (33, 4, 5)])
def test_emulated_four_eqn_tiny_patches_and_push(N, oracle_lib):
    desc, U = _case(3, N, seed=3)
    Q = pb.pad_periodic(U)
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, 1.0e-3)
    Uo = oracle_lib.advance_stage(desc, [1.0], [1.0], [Q], [Fo], [So])
    Ue = emu_host.fused_stage(desc, [1.0], [1.0], [Q], 1.0e-3, math=0, push=True)
    inner = (slice(None),) + (slice(4, -4),) * 3
    assert np.array_equal(Ue[inner], Uo[inner])
    assert np.array_equal(Ue, pb.pad_periodic(np.ascontiguousarray(Ue[inner])))
