"""The reference's own acceptance tests, run on the oracle: L2 convergence rate > 4.8 at the finest grid pair
for WCNS5_JS_HLLC_HLL (tests/{2D,3D}_convergence_test_single_species/convergence_test.py:13-25, 196-281 and
tests/3D_convergence_test_five_eqn_allaire/convergence_test.py; 8^d patches, dt = 0.001*(2/8)/2^L,
num_steps_base*2^L steps with num_steps_base = 8 in 2D and 1 in 3D)."""
import numpy as np
import pytest

from hamers_b200 import problems as pb

EXPECTED_RATE = 4.8


def _run(orc, dim, model, levels, steps_base, scheme=0):
    errs = []
    for L in range(levels):
        N = 8 * 2 ** L
        if model == 0:
            U, dx, gam = pb.convergence_single_species(dim, N)
        else:
            U, dx, gam = pb.convergence_five_eqn(dim, N)
        lvl = orc.PatchDesc(dim=dim, n=(N,) * dim, model=model, ns=len(gam), gamma=gam, dx=dx, scheme=scheme)
        dt = 0.001 * (2.0 / 8) / 2 ** L
        nsteps = steps_base * 2 ** L
        orc.level_advance(lvl, (8,) * dim, U, dt, nsteps, nthreads=0)
        t = dt * nsteps
        if model == 0:
            errs.append(pb.error_norms(U[0], pb.exact_density_single_species(dim, N, t), dx))
        else:
            errs.append(pb.error_norms(U[-2], pb.exact_Z1_five_eqn(dim, N, t), dx))
    rates = [np.log2(errs[i - 1][1] / errs[i][1]) for i in range(1, levels)]
    return errs, rates


def test_2d_single_species_order(oracle_lib):
    errs, rates = _run(oracle_lib, 2, 0, 4, 8)
    assert rates[-1] > EXPECTED_RATE, (errs, rates)


def test_3d_single_species_order(oracle_lib):
    errs, rates = _run(oracle_lib, 3, 0, 4, 1)
    assert rates[-1] > EXPECTED_RATE, (errs, rates)


def test_3d_five_eqn_order(oracle_lib):
    errs, rates = _run(oracle_lib, 3, 1, 4, 1)
    assert rates[-1] > EXPECTED_RATE, (errs, rates)


def test_2d_five_eqn_order(oracle_lib):
    errs, rates = _run(oracle_lib, 2, 1, 4, 8)
    assert rates[-1] > EXPECTED_RATE, (errs, rates)


@pytest.mark.parametrize("scheme,rate", [(1, 4.8), (2, 5.8)])
def test_2d_single_species_order_of_the_other_interpolators(scheme, rate, oracle_lib):
    """WCNS5_Z_HLLC_HLL must exceed 4.8 and WCNS6_LD_HLLC_HLL 5.8
    (tests/2D_convergence_test_single_species/convergence_test.py:8-16)."""
    errs, rates = _run(oracle_lib, 2, 0, 4, 8, scheme=scheme)
    assert rates[-1] > rate, (errs, rates)


def test_level_advance_is_patch_size_independent(oracle_lib):
    """Tiling the level into patches must not change the answer (ghost fill = plain periodic copies)."""
    U, dx, gam = pb.random_state(2, (32, 24), seed=3, shock=True)
    lvl = oracle_lib.PatchDesc(dim=2, n=(32, 24), gamma=gam, dx=dx)
    A, B = U.copy(), U.copy()
    oracle_lib.level_advance(lvl, (8, 8), A, 1e-4, 2)
    oracle_lib.level_advance(lvl, (32, 24), B, 1e-4, 2)
    assert np.array_equal(A, B)
