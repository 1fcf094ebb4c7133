"""SURVEY row f3 on the GPU: the two-level hierarchy of hamers_b200/amr.py through the C ABI (every operator a kernel of
libhamers_b200.so) against the oracle's independent restatement (oracle/amr.py), BIT FOR BIT in the reference-order build:
single-species and four-eqn conservative models, periodic and FLOW boundaries, fine patches that touch the boundary or span
a periodic direction; composite-grid conservation; a 3-D hierarchy; and BASELINE.json's config 4 as an adapted problem --
the 2-D Richtmyer-Meshkov deck's state (SF6 / air, WCNS6_LD_HLLC_HLL, FLOW in x, periodic in y) on a static two-level
hierarchy (SAMRAI's regridding is out of scope)."""
import numpy as np
import pytest

from hamers_b200 import problems as pb

pytestmark = pytest.mark.gpu
G = 4


@pytest.mark.parametrize("model,scheme", [(0, 0), (2, 2)])
@pytest.mark.parametrize("clo,chi,periodic", [((4, 4), (12, 12), (True, True)), ((4, 0), (12, 16), (False, True)),
                                             ((0, 0), (10, 16), (False, True))])
def test_two_level_steps_match_the_oracle_hierarchy(clo, chi, periodic, model, scheme, oracle_lib, product_lib):
    import torch
    from test_amr_emulated import build_pair

    H, O = build_pair(model, 16, clo, chi, periodic, {}, scheme)
    inner = (slice(None),) + (slice(G, -G),) * 2
    t0 = H.composite_totals()
    dt = 0.1 * O.dc.dx[0]
    for _ in range(2):
        H.advance(dt)
        O.advance(dt)
        torch.cuda.synchronize()
        assert np.array_equal(H.Uc.cpu().numpy()[inner], O.Uc[inner])
        assert np.array_equal(H.Uf.cpu().numpy()[inner], O.Uf[inner])
    if all(periodic):
        assert (np.abs(H.composite_totals() - t0) <= 4e-15 * np.abs(t0)).all()
    H.close()


def test_three_dimensional_hierarchy(oracle_lib, product_lib):
    import torch
    from hamers_b200 import abi
    from hamers_b200.amr import TwoLevelHierarchy
    from oracle import amr

    N, clo, chi = 12, (2, 0, 3), (9, 12, 10)
    U, dx, gam = pb.convergence_single_species(3, N)
    Uf, _, _ = pb.convergence_single_species(3, 2 * N)
    desc = oracle_lib.PatchDesc(dim=3, n=(N,) * 3, model=0, ns=1, gamma=gam, dx=dx)
    O = amr.TwoLevelOracle(desc, clo, chi, 2, (True, True, True))
    H = TwoLevelHierarchy(3, (N,) * 3, clo, chi, 2, (True, True, True), species_gamma=gam, dx=dx, math=abi.MATH_EXACT)
    box = (slice(None),) + tuple(slice(2 * clo[a], 2 * chi[a]) for a in (2, 1, 0))
    inner = (slice(None),) + (slice(G, -G),) * 3
    O.Uc[inner], O.Uf[inner] = U, Uf[box]
    O.Uc[amr._sl(3, clo, chi, G)] = amr.conservative_coarsen(O.Uf[inner], 3, O.r, desc.dx, O.df.dx)
    H.set_coarse(U)
    H.set_fine(np.ascontiguousarray(Uf[box]))
    H.coarsen_fine_onto_coarse()
    t0 = H.composite_totals()
    H.advance(0.2 * dx[0])
    O.advance(0.2 * dx[0])
    torch.cuda.synchronize()
    assert np.array_equal(H.Uc.cpu().numpy()[inner], O.Uc[inner]) and np.array_equal(H.Uf.cpu().numpy()[inner], O.Uf[inner])
    assert (np.abs(H.composite_totals() - t0) <= 4e-15 * np.abs(t0)).all()
    H.close()


def test_config_4_richtmyer_meshkov_on_a_static_two_level_hierarchy(oracle_lib, product_lib):
    """problems/support_files/2D_Richtmyer_Meshkov_instability: FOUR_EQN_CONSERVATIVE, species of the deck, WCNS6_LD_HLLC_HLL,
    FLOW boundaries in x, periodic in y, ratio 2; the fine patch covers the shock and the perturbed interface.  dt from the
    level's stable dt at CFL 0.5 like the deck.  Bit-identical to the oracle hierarchy."""
    import torch
    from hamers_b200 import abi
    from hamers_b200.amr import TwoLevelHierarchy
    from oracle import amr

    N = (256, 16)                                         # coarse level of the adapted box [0, 0.004] x [0, 0.0005]
    x_up = (0.004, 0.0005)
    U, dx, gam, R = pb.richtmyer_meshkov_2d(N, x_up)
    clo, chi = (8, 0), (72, 16)                           # interface at x ~ 0.4 mm, shock at 0.7 mm: coarse cells 19..51
    Uf_all, dxf, _, _ = pb.richtmyer_meshkov_2d((2 * N[0], 2 * N[1]), x_up)
    desc = oracle_lib.PatchDesc(dim=2, n=N, model=2, ns=2, gamma=gam, R=R, dx=dx, scheme=2)
    periodic = (False, True)
    O = amr.TwoLevelOracle(desc, clo, chi, 2, periodic)
    H = TwoLevelHierarchy(2, N, clo, chi, 2, periodic, flow_model=abi.FOUR_EQN_CONSERVATIVE, species_gamma=gam, species_R=R, dx=dx,
                          math=abi.MATH_EXACT, scheme=abi.WCNS6_LD)
    box = (slice(None),) + tuple(slice(2 * clo[a], 2 * chi[a]) for a in (1, 0))
    inner = (slice(None),) + (slice(G, -G),) * 2
    O.Uc[inner], O.Uf[inner] = U, Uf_all[box]
    O.Uc[amr._sl(2, clo, chi, G)] = amr.conservative_coarsen(O.Uf[inner], 2, O.r, desc.dx, O.df.dx)
    H.set_coarse(U)
    H.set_fine(np.ascontiguousarray(Uf_all[box]))
    H.coarsen_fine_onto_coarse()
    # stable dt of the coarse level at CFL 0.5 (Euler::computeSpectralRadiusesAndStableDtOnPatch)
    sr = torch.zeros(4, dtype=torch.float64, device="cuda")
    H._fill_coarse(H.Uc)
    H.coarse.plan.max_wave_speed(H.Uc, sr)
    dt = 0.5 / (float(sr[3]) + 1.0e-15)
    for _ in range(3):
        H.advance(dt)
        O.advance(dt)
    torch.cuda.synchronize()
    got_c, got_f = H.Uc.cpu().numpy()[inner], H.Uf.cpu().numpy()[inner]
    assert np.isfinite(got_c).all() and got_c[:2].min() > -1e-6            # partial densities stay non-negative to round-off
    assert np.array_equal(got_c, O.Uc[inner]) and np.array_equal(got_f, O.Uf[inner])
    # both species stream through the FLOW boundaries, so their totals change -- by the same bits as the oracle's
    tot = O.composite_totals()
    assert np.allclose(H.composite_totals(), tot, rtol=1e-13, atol=1e-13 * np.abs(tot).max())
    H.close()
