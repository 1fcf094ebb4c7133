"""The oracle at THREE species against the reference's own code: the HLLC / HLLC-HLL point kernels of the five-eqn and
four-eqn flow models (`static inline` functions that take d_num_species as an argument, compiled verbatim by
oracle/build_ref.py) -- committed outputs in tests/golden/three_species_kernels.npz (generator:
tests/golden/make_golden_three_species.py) and, when oracle/_ref is present, live on fresh inputs.  The mixture chains in
front of them are the loops pinned at two species (tests/test_oracle_pinned.py, test_oracle_four_eqn.py) run over one more
species; their closure is checked here against numpy restatements of the mixture rules."""
import ctypes as C
import os

import numpy as np
import pytest

from hamers_b200 import problems as pb
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "three_species_kernels.npz"))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libhamers_ref.so")
FE, FC = orc.FIVE_EQN_ALLAIRE, orc.FOUR_EQN_CONSERVATIVE


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("tag,model", [("fe", FE), ("fc", FC)])
def test_riemann_kernels_at_three_species_match_the_reference(tag, model, dim, oracle_lib):
    gam = tuple(GOLD["gamma_fe"]) if model == FE else tuple(GOLD["gamma_R_fc"])
    differs = 0
    for d in range(dim):
        key = f"rp_{tag}{dim}d{d}"
        VL, VR = GOLD[key + "_VL"], GOLD[key + "_VR"]
        for n in range(VL.shape[0]):
            F1, F2, vm = oracle_lib.riemann_point(model, dim, 3, gam, d, VL[n], VR[n])
            tl, tr = oracle_lib.side_thermo(model, dim, 3, gam, VL[n]), oracle_lib.side_thermo(model, dim, 3, gam, VR[n])
            assert np.array_equal(np.array([tl[0], tr[0], tl[1], tr[1], tl[2], tr[2]]), GOLD[key + "_thermo"][n])
            assert np.array_equal(F1, GOLD[key + "_F_HLLC"][n]), (key, n, "HLLC")
            assert np.array_equal(F2, GOLD[key + "_F_HYB"][n]), (key, n, "HLLC-HLL")
            if model == FE:
                assert vm == GOLD[key + "_vel_mid"][n], (key, n, "midpoint velocity")
            differs += int(not np.array_equal(F1, F2))
    assert differs > 100          # the hybrid flux really differs from HLLC on most faces


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("model", [FE, FC])
def test_riemann_kernels_at_three_species_match_the_reference_live(model, oracle_lib):
    import sys

    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden as mg
    import make_golden_four_eqn as mg4

    lib = C.CDLL(REF_SO)
    lib.ref_riemann_point.restype = C.c_int
    lib.ref_riemann_point_fc.restype = C.c_int
    rng = np.random.default_rng(77)
    gam = (1.5, 1.3, 1.67) if model == FE else (1.5, 1.3, 1.67, 2.0, 0.5, 1.1)
    for dim in (2, 3):
        for d in range(dim):
            if model == FE:
                VL, VR = mg.riemann_inputs(rng, 1, dim, 3, n=60)
                VL[:, 3 + dim + 1:] *= 0.5
                VR[:, 3 + dim + 1:] *= 0.5
            else:
                VL, VR = mg4.riemann_inputs(rng, dim, 3, n=60)
            for a, b in zip(VL, VR):
                F1, F2, vm = oracle_lib.riemann_point(model, dim, 3, gam, d, a, b)
                if model == FE:
                    R1, R2, rvm, _ = mg.ref_riemann(lib, 1, dim, 3, gam, d, a, b)
                    assert vm == rvm
                else:
                    tl, tr = oracle_lib.side_thermo(FC, dim, 3, gam, a), oracle_lib.side_thermo(FC, dim, 3, gam, b)
                    R1, R2 = mg4.ref_riemann(lib, dim, 3, d, a, b, (tl[0], tr[0], tl[1], tr[1], tl[2], tr[2]))
                assert np.array_equal(F1, R1) and np.array_equal(F2, R2)


def test_three_species_mixture_closures(oracle_lib):
    """The side state the oracle rebuilds in front of the Riemann kernels obeys the mixture rules at three species:
    five-eqn: rho = sum Z_i rho_i, 1/(gamma - 1) = sum Z_i/(gamma_i - 1) with Z_3 = 1 - Z_1 - Z_2, c^2 = gamma p/rho;
    four-eqn: gamma = sum Y_i c_p,i / sum Y_i c_v,i, c^2 = gamma p/rho."""
    rng = np.random.default_rng(8)
    gam, R = (1.6, 1.4, 1.25), (0.7, 1.3, 1.0)
    for _ in range(200):
        zr = rng.uniform(0.1, 2.0, 3)
        vel = rng.uniform(-2, 2, 3)
        p = rng.uniform(0.3, 5.0)
        Z = rng.uniform(0.05, 0.45, 2)
        V = np.concatenate([zr, vel, [p], Z])
        rho, c, eps = oracle_lib.side_thermo(FE, 3, 3, gam, V)
        Zall = np.array([Z[0], Z[1], 1.0 - Z[0] - Z[1]])
        g = 1.0 / np.sum(Zall / (np.array(gam) - 1.0)) + 1.0
        assert np.isclose(rho, zr.sum(), rtol=1e-15) and np.isclose(c * c, g * p / rho, rtol=1e-13)
        assert np.isclose(eps, p / ((g - 1.0) * rho), rtol=1e-13)
        V4 = np.concatenate([zr, vel, [p]])
        rho4, c4, eps4 = oracle_lib.side_thermo(FC, 3, 3, gam + R, V4)
        Y = zr / zr.sum()
        g4 = pb.mixture_gamma_mass_fractions(Y, gam, R)
        assert np.isclose(rho4, zr.sum(), rtol=1e-15) and np.isclose(c4 * c4, g4 * p / rho4, rtol=1e-13)


@pytest.mark.parametrize("model", [FE, FC])
def test_three_species_stage_conserves(model, oracle_lib):
    """One forward-Euler stage of the oracle on a periodic box: every conservative equation keeps its sum to round-off."""
    if model == FE:
        U, dx, gam = pb.random_state_three_species(2, (20, 16), FE, seed=4, shock=False)
        desc = oracle_lib.PatchDesc(dim=2, n=(20, 16), model=FE, ns=3, gamma=gam, dx=dx)
        ncons = 3 + 2 + 1
    else:
        U, dx, gam, R = pb.random_state_three_species(2, (20, 16), FC, seed=4, shock=False)
        desc = oracle_lib.PatchDesc(dim=2, n=(20, 16), model=FC, ns=3, gamma=gam, R=R, dx=dx)
        ncons = desc.neq
    Q = pb.pad_periodic(U)
    F, S = oracle_lib.compute_flux_and_source(desc, Q, 1.0e-3)
    Un = oracle_lib.advance_stage(desc, [1.0], [1.0], [Q], [F], [S])
    inner = (slice(None),) + (slice(4, -4),) * 2
    for e in range(ncons):
        assert abs(Un[inner][e].sum() - U[e].sum()) <= 1.0e-12 * np.abs(U[e]).sum(), e
    assert np.isfinite(Un[inner]).all()
