"""SURVEY row f3, oracle side: the FOUR_EQN_CONSERVATIVE flow model.

Pins: the oracle's mixture chain (rho, Y_i, epsilon, mixture c_p / c_v / gamma from the mass fractions, p, Psi_i, sound
speed), the (rho, c, epsilon) of an interpolated side, the bounds flag, the face averages, the characteristic projection and
its inverse, and the HLLC / HLLC-HLL point kernels are compared BIT FOR BIT with the reference's own code for this model
(FlowModelFourEqnConservative.cpp, FlowModelBasicUtilitiesFourEqnConservative.cpp, EquationOfStateMixingRulesIdealGas.cpp,
Riemann_solvers/FlowModelRiemannSolverFourEqnConservativeHLLC{,-HLL}.cpp; compiled by oracle/build_ref.py into oracle/_ref):
committed outputs in tests/golden/four_eqn_kernels.npz (generator: tests/golden/make_golden_four_eqn.py) and, when
oracle/_ref is present, live on fresh inputs.  The reference ships no test of its own for this model; the acceptance
protocol of its convergence tests (L2 rate > 4.8 at the finest grid pair) is applied to a two-species advection problem."""
import ctypes as C
import os

import numpy as np
import pytest

from hamers_b200 import problems as pb
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "four_eqn_kernels.npz"))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libhamers_ref.so")
FC = orc.FOUR_EQN_CONSERVATIVE


def test_mixture_chain_side_state_and_bounds_match_reference_golden(oracle_lib):
    got = np.array([oracle_lib.path_points7(v) for v in GOLD["pp7_in"]])
    ref = GOLD["pp7_out"]
    assert np.array_equal(got, ref, equal_nan=True)
    assert 0.2 < ref[:, 14].mean() < 0.9          # both outcomes of the bounds flag are well represented
    # the chain is thermodynamically consistent: c^2 = gamma p / rho for a mixture of ideal gases
    ok = np.isfinite(ref[:, 10])              # a few inputs have negative internal energy (NaN sound speed on both sides)
    assert ok.sum() > 250
    assert np.allclose(ref[ok, 10] ** 2, (ref[:, 6] * ref[:, 7] / ref[:, 0])[ok], rtol=1e-13)


def test_projection_and_back_projection_match_reference_golden(oracle_lib):
    got = np.array([oracle_lib.path_points8(v) for v in GOLD["pp8_in"]])
    assert np.array_equal(got, GOLD["pp8_out"])


@pytest.mark.parametrize("dim", [2, 3])
def test_riemann_point_kernels_match_reference_golden(dim, oracle_lib):
    gr = tuple(GOLD["gamma_R"])
    differs = 0
    for d in range(dim):
        key = f"rp_fc{dim}d{d}"
        VL, VR = GOLD[key + "_VL"], GOLD[key + "_VR"]
        for n in range(VL.shape[0]):
            F1, F2, _ = oracle_lib.riemann_point(FC, dim, 2, gr, d, VL[n], VR[n])
            tl, tr = oracle_lib.side_thermo(FC, dim, 2, gr, VL[n]), oracle_lib.side_thermo(FC, dim, 2, gr, VR[n])
            assert np.array_equal(np.array([tl[0], tr[0], tl[1], tr[1], tl[2], tr[2]]), GOLD[key + "_thermo"][n])
            assert np.array_equal(F1, GOLD[key + "_F_HLLC"][n]), (key, n, "HLLC")
            assert np.array_equal(F2, GOLD[key + "_F_HYB"][n]), (key, n, "HLLC-HLL")
            differs += int(not np.array_equal(F1, F2))
    assert differs > 100          # the hybrid flux really differs from HLLC on most faces


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference)")
def test_mixture_chain_matches_reference_live(oracle_lib):
    lib = C.CDLL(REF_SO)
    rng = np.random.default_rng(5)
    for _ in range(500):
        v = np.concatenate([rng.uniform(0.05, 2.0, 2), rng.uniform(-2, 2, 3), rng.uniform(4, 12, 1), rng.uniform(1.1, 1.9, 2),
                            rng.uniform(0.3, 3.0, 2), rng.uniform(-0.05, 2.0, 2), rng.uniform(-2, 2, 3), rng.uniform(-0.2, 5.0, 1)])
        out = (C.c_double * 15)()
        lib.ref_path_points7((C.c_double * 16)(*v), out)
        assert np.array_equal(np.array(oracle_lib.path_points7(v)), np.array(out[:]), equal_nan=True)


def test_single_species_limit(oracle_lib):
    """Two species with equal gamma and R at uniform composition are one ideal gas: the fluxes of the four-eqn model equal
    the single-species model's (partial densities summed) up to the epsilon = 1e-15 of the nonlinear weights."""
    N = (12, 10)
    U, dx, gam = pb.random_state(2, N, model=0, seed=8, shock=True)
    Y0 = 0.5        # uniform composition: the partial densities are exact multiples of rho (the interpolation is nonlinear)
    U4 = np.stack([U[0] * Y0, U[0] * (1.0 - Y0), U[1], U[2], U[3]])
    d1 = oracle_lib.PatchDesc(dim=2, n=N, model=0, ns=1, gamma=gam, dx=dx)
    d4 = oracle_lib.PatchDesc(dim=2, n=N, model=FC, ns=2, gamma=(gam[0], gam[0]), R=(1.0, 1.0), dx=dx)
    F1, _ = oracle_lib.compute_flux_and_source(d1, pb.pad_periodic(U), 1e-3)
    F4, _ = oracle_lib.compute_flux_and_source(d4, pb.pad_periodic(U4), 1e-3)
    for a in range(2):
        scale = np.abs(F1[a]).max()
        assert np.abs(F4[a][0] + F4[a][1] - F1[a][0]).max() <= 1e-10 * scale
        assert np.abs(F4[a][2:] - F1[a][1:]).max() <= 1e-10 * scale


def test_convergence_order_two_species(oracle_lib):
    """The protocol of tests/2D_convergence_test_*/convergence_test.py (N = 8..64, dt = 0.001*(2/8)/2^L, 8*2^L steps, L2 rate
    at the finest pair > 4.8) on the two-species advection problem of pb.convergence_four_eqn (exact solution: the mass
    fraction wave translated with the uniform velocity)."""
    errs = []
    for L in range(4):
        N = 8 * 2 ** L
        U, dx, gam, R = pb.convergence_four_eqn(2, N)
        lvl = oracle_lib.PatchDesc(dim=2, n=(N, N), model=FC, ns=2, gamma=gam, R=R, dx=dx)
        dt = 0.001 * (2.0 / 8) / 2 ** L
        nsteps = 8 * 2 ** L
        oracle_lib.level_advance(lvl, (8, 8), U, dt, nsteps, nthreads=0)
        errs.append(pb.error_norms(U[0], pb.exact_rhoY1_four_eqn(2, N, dt * nsteps), dx))
    rate = np.log2(errs[-2][1] / errs[-1][1])
    assert rate > 4.8, (errs, rate)
