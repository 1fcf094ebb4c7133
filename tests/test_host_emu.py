"""CPU check of the CUDA kernels' indexing and arithmetic: tests/host_emu compiles the very same
`__host__ __device__` thread functions the sm_100a kernels call (hamers_b200/csrc/hb2_core.cuh) with g++ and
drives them from loops standing in for the CUDA grid.  Compared against the oracle: bit-identical for the
exact-arithmetic variant, <= 1e-12 for the fast variant.  (The GPU parity tests remain the parity tests proper;
this catches mistakes before GPU time is spent.  The emulation is never linked into the product.)"""
import numpy as np
import pytest

import emu_host
from common import CASES, RTOL, assert_fast_parity, flux_source_spread, interior, make_case, rel_err
from hamers_b200 import problems as pb

SSPRK3_ALPHA = [[1.0], [3.0 / 4.0, 1.0 / 4.0], [1.0 / 3.0, 0.0, 2.0 / 3.0]]
SSPRK3_BETA = [[1.0], [0.0, 1.0 / 4.0], [0.0, 0.0, 2.0 / 3.0]]


@pytest.mark.parametrize("kind", ["random", "smooth"])
@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("name", list(CASES))
def test_emulated_flux_and_source(name, math, kind, oracle_lib):
    desc, U = make_case(name, kind)
    Q = pb.pad_periodic(U)
    dt = 1.0e-3
    S0 = np.random.default_rng(7).standard_normal((desc.neq,) + desc.cell_shape)
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt, source=S0.copy())
    Fe, Se = emu_host.flux_and_source(desc, Q, dt, math=math, source=S0.copy())
    # fast build: every entry beyond 1e-12 must be explained by the oracle's own conditioning at that entry (a branch within
    # an ulp of its threshold, an ill-conditioned face): tests/common.py OracleSpread
    sp = flux_source_spread(oracle_lib, desc, U, dt, S0) if math == 1 else None
    for a in range(desc.dim):
        assert not np.isnan(Fe[a]).any()
        if math == 0:
            assert np.array_equal(Fe[a], Fo[a]), f"dir {a}: max diff {np.abs(Fe[a] - Fo[a]).max()}"
        else:
            assert_fast_parity(Fe[a], Fo[a], f"dir {a}", spread=sp.spread[a])
    if math == 0:
        assert np.array_equal(Se, So)
    else:
        assert_fast_parity(Se, So, "source", spread=sp.spread[desc.dim])


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("name", list(CASES))
def test_emulated_fused_stages(name, math, oracle_lib):
    desc, U = make_case(name, "random")
    dt = 5.0e-4
    states = [pb.pad_periodic(U)]
    for sn in range(3):
        F, S = oracle_lib.compute_flux_and_source(desc, states[-1], dt)
        n = sn + 1
        Uo = oracle_lib.advance_stage(desc, SSPRK3_ALPHA[sn], SSPRK3_BETA[sn], states, [None] * (n - 1) + [F], [None] * (n - 1) + [S])
        Ue = emu_host.fused_stage(desc, SSPRK3_ALPHA[sn], SSPRK3_BETA[sn], states, dt, math=math)
        if math == 0:
            assert np.array_equal(interior(desc, Ue), interior(desc, Uo)), f"stage {sn}"
        else:
            assert rel_err(interior(desc, Ue), interior(desc, Uo)) <= RTOL, f"stage {sn}"
        states.append(pb.pad_periodic(np.ascontiguousarray(interior(desc, Uo))))


@pytest.mark.parametrize("seg_len,bx", [(5, 4), (3, 7), (0, 0)])
def test_emulated_tiling_is_invisible(seg_len, bx, oracle_lib):
    """Segment lengths (sweeps: seg_len, sensor pass: bx) are launch parameters only: results must not depend on them."""
    desc, U = make_case("ss3d", "random")
    Q = pb.pad_periodic(U)
    Fo, _ = oracle_lib.compute_flux_and_source(desc, Q, 1e-3)
    Fe, _ = emu_host.flux_and_source(desc, Q, 1e-3, math=0, bx=bx, seg_len=seg_len)
    for a in range(3):
        assert np.array_equal(Fe[a], Fo[a])


@pytest.mark.parametrize("model,dim,N,seg_len", [(0, 2, (150, 40), 0), (0, 2, (150, 40), 70), (1, 2, (135, 37), 0),
                                                 (0, 3, (36, 40, 34), 0), (1, 3, (34, 33, 40), 0)])
def test_emulated_ring_wraparound(model, dim, N, seg_len, oracle_lib):
    """Pencils longer than the shared-memory rings (32 slots in y/z, 128 in x): the rings wrap several times while a
    block marches, so a slot reused too early (or a stale mirrored slot) changes the result."""
    U, dx, gam = pb.random_state(dim, N, model=model, seed=17, shock=True)
    desc = oracle_lib.PatchDesc(dim=dim, n=N, model=model, ns=len(gam), gamma=gam, dx=dx)
    Q = pb.pad_periodic(U)
    dt = 1.0e-3
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt)
    Fe, Se = emu_host.flux_and_source(desc, Q, dt, math=0, seg_len=seg_len)
    for a in range(dim):
        assert np.array_equal(Fe[a], Fo[a]), f"dir {a}"
    assert np.array_equal(Se, So)
    Uo = oracle_lib.advance_stage(desc, [1.0], [1.0], [Q], [Fo], [So])
    for math in (0, 1):
        Ue = emu_host.fused_stage(desc, [1.0], [1.0], [Q], dt, math=math, seg_len=seg_len)
        if math == 0:
            assert np.array_equal(interior(desc, Ue), interior(desc, Uo))
        else:
            assert_fast_parity(interior(desc, Ue), interior(desc, Uo))


@pytest.mark.parametrize("model,dim,N", [(0, 3, (64, 6, 50)), (1, 2, (64, 70)), (0, 2, (100, 32))])
def test_emulated_steady_iterations(model, dim, N, oracle_lib):
    """Blocks that have all their pencils run the interior iterations of a march through the STEADY variant of the iteration
    body (no per-phase "wanted" tests, hb2_sweep.cuh: pipeline_step); blocks with missing pencils and the first / last
    iterations run the general one.  All three SSP-RK3 rows with the fused ghost push, both arithmetic variants."""
    U, dx, gam = pb.random_state(dim, N, model=model, seed=23, shock=True)
    desc = oracle_lib.PatchDesc(dim=dim, n=N, model=model, ns=len(gam), gamma=gam, dx=dx)
    Q = pb.pad_periodic(U)
    dt = 1.0e-3
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt)
    Fe, Se = emu_host.flux_and_source(desc, Q, dt, math=0)
    for a in range(dim):
        assert np.array_equal(Fe[a], Fo[a]), f"dir {a}"
    assert np.array_equal(Se, So)
    for alpha, beta in (([1.0], [1.0]), ([0.75, 0.25], [0.0, 0.25]), ([1.0 / 3.0, 0.0, 2.0 / 3.0], [0.0, 0.0, 2.0 / 3.0])):
        m = len(alpha)
        older = []
        for k in range(m - 1):
            V = U * (1.0 + 0.01 * k)
            if model == 1:
                V[-2:] = U[-2:]
            older.append(pb.pad_periodic(V))
        states = older + [Q]
        none = [None] * (m - 1)
        Uo = oracle_lib.advance_stage(desc, alpha, beta, states, none + [Fo], none + [So])
        for math in (0, 1):
            Ue = emu_host.fused_stage(desc, alpha, beta, states, dt, math=math, push=True)
            if math == 0:
                assert np.array_equal(interior(desc, Ue), interior(desc, Uo))
            else:
                assert_fast_parity(interior(desc, Ue), interior(desc, Uo))
            assert np.array_equal(Ue, pb.pad_periodic(np.ascontiguousarray(interior(desc, Ue))))


@pytest.mark.parametrize("name", ["ss2d", "ss3d", "fe3d"])
def test_emulated_push_fills_the_ghosts(name, oracle_lib):
    """Ghost fill fused into the update (push_cell): with the patch as its own periodic neighbour every ghost cell of
    the new state -- faces, edges, corners -- must be the periodic image of the new interior, and the interior must
    be what the stage without push produces."""
    desc, U = make_case(name, "random")
    Q = pb.pad_periodic(U)
    plain = emu_host.fused_stage(desc, [1.0], [1.0], [Q], 5.0e-4, math=0)
    pushed = emu_host.fused_stage(desc, [1.0], [1.0], [Q], 5.0e-4, math=0, push=True)
    assert np.array_equal(interior(desc, pushed), interior(desc, plain))
    assert np.array_equal(pushed, pb.pad_periodic(np.ascontiguousarray(interior(desc, pushed))))


@pytest.mark.parametrize("scheme", [1, 2])
@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("name", ["ss2d", "fe3d"])
def test_emulated_other_interpolators(name, math, scheme, oracle_lib):
    """WCNS5-Z / WCNS6-LD through the emulated kernels (compiled with -DHB2_SCHEME): reference-order arithmetic
    bit-identical to the oracle, re-associated arithmetic within the fast-build criterion (WCNS6-LD with its larger
    share of ill-conditioned faces, tests/test_oracle_conditioning.py)."""
    import dataclasses

    desc, U = make_case(name, "random")
    desc = dataclasses.replace(desc, scheme=scheme)
    Q = pb.pad_periodic(U)
    dt = 1.0e-3
    S0 = np.random.default_rng(7).standard_normal((desc.neq,) + desc.cell_shape)
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt, source=S0.copy())
    Fe, Se = emu_host.flux_and_source(desc, Q, dt, math=math, source=S0.copy())
    F1, S1 = oracle_lib.compute_flux_and_source(desc, Q, dt)
    Uo = oracle_lib.advance_stage(desc, [1.0], [1.0], [Q], [F1], [S1])
    Ue = emu_host.fused_stage(desc, [1.0], [1.0], [Q], dt, math=math)
    frac = 5.0e-3 if scheme == 2 else 1.0e-3
    for a in range(desc.dim):
        if math == 0:
            assert np.array_equal(Fe[a], Fo[a]), f"dir {a}"
        else:
            assert_fast_parity(Fe[a], Fo[a], f"dir {a}", frac)
    if math == 0:
        assert np.array_equal(Se, So) and np.array_equal(interior(desc, Ue), interior(desc, Uo))
    else:
        assert_fast_parity(Se, So, "source", frac)
        assert_fast_parity(interior(desc, Ue), interior(desc, Uo), "fused stage", frac)


@pytest.mark.parametrize("model,ns", [(0, 1), (1, 2)])
@pytest.mark.parametrize("N", [(4, 4, 4), (5, 4, 6), (4, 33, 5), (65, 4, 4), (5, 5, 6), (7, 7, 7)])
def test_emulated_tiny_and_ragged_patches(N, model, ns, oracle_lib):
    """Patches as narrow as the ghost width (every cell is near BOTH faces of a direction: the fused ghost push then
    stores it into both neighbours), pencils shorter than one chunk, odd extents: fluxes, the fused stage and the pushed
    ghosts against the oracle."""
    U, dx, gam = pb.random_state(3, N, model=model, seed=3, shock=True)
    desc = oracle_lib.PatchDesc(dim=3, n=N, model=model, ns=ns, gamma=gam, dx=dx)
    Q = pb.pad_periodic(U)
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, 1.0e-3)
    Uo = oracle_lib.advance_stage(desc, [1.0], [1.0], [Q], [Fo], [So])
    Fe, _ = emu_host.flux_and_source(desc, Q, 1.0e-3, math=0)
    Ue = emu_host.fused_stage(desc, [1.0], [1.0], [Q], 1.0e-3, math=0, push=True)
    for a in range(3):
        assert np.array_equal(Fe[a], Fo[a]), f"dir {a}"
    assert np.array_equal(interior(desc, Ue), interior(desc, Uo))
    assert np.array_equal(Ue, pb.pad_periodic(np.ascontiguousarray(interior(desc, Ue))))
    Uf = emu_host.fused_stage(desc, [1.0], [1.0], [Q], 1.0e-3, math=1)
    assert_fast_parity(interior(desc, Uf), interior(desc, Uo))


@pytest.mark.parametrize("name", ["fe2d", "fe3d"])
def test_emulated_bounds_check_follows_the_reference_per_direction(name, oracle_lib):
    """The reference's five-eqn c^2 check accumulates Y_i Psi_i only in its x-direction block; its y / z blocks test
    Gamma p / rho > 0 (FlowModelBasicUtilitiesFiveEqnAllaire.cpp:6654-6678 vs 6968-6992, 7281-7305; pinned in
    tests/test_oracle_pinned.py).  With species gammas (1.0005, 3) an interpolated volume fraction just below zero gives
    -1 < Gamma < 0, where the two forms disagree, so the first-order fallback is taken on different faces per
    direction: the emulated kernels must follow the oracle bit for bit there too."""
    import dataclasses

    desc, U = make_case(name, "random")
    desc = dataclasses.replace(desc, gamma=(1.0005, 3.0))
    # the state exercises the disagreement: an interpolated side as it occurs here, flagged differently by direction
    side = [0.3, 0.4, 0.1, -0.2, 0.3, 1.0, -0.002, 1.0005, 3.0]
    assert oracle_lib.path_points5(side + [0.0])[0] == 1.0 and oracle_lib.path_points5(side + [1.0])[0] == 0.0
    Q = pb.pad_periodic(U)
    dt = 1.0e-3
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt)
    Fe, Se = emu_host.flux_and_source(desc, Q, dt, math=0)
    for a in range(desc.dim):
        assert np.array_equal(Fe[a], Fo[a]), f"dir {a}"
    assert np.array_equal(Se, So)
    # re-associated arithmetic: same faces fall back (a wrong flag shows as an O(1) flux difference)
    Ff, Sf = emu_host.flux_and_source(desc, Q, dt, math=1)
    for a in range(desc.dim):
        assert_fast_parity(Ff[a], Fo[a], f"dir {a}", 5.0e-3)


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("name", ["ss2d", "ss3d", "fe3d"])
def test_emulated_sweeps_on_a_six_ghost_layout(name, math, oracle_lib):
    """The Navier-Stokes application allocates the state with six ghost cells (max of the convective 4 and the diffusive
    6); the convective kernels address cells through the layout's ghost width and strides only, so the same kernels on a
    six-ghost array give the results of the four-ghost one: fluxes / sources bit for bit, the fused stage on the interior."""
    desc, U = make_case(name, "random")
    Q4, Q6 = pb.pad_periodic(U, 4), pb.pad_periodic(U, 6)
    dt = 1.0e-3
    F4, S4 = emu_host.flux_and_source(desc, Q4, dt, math=math)
    F6, S6 = emu_host.flux_and_source(desc, Q6, dt, math=math, ghosts=6)
    for a in range(desc.dim):
        assert np.array_equal(F4[a], F6[a]), f"dir {a}"
    assert np.array_equal(S4, S6)
    U4 = emu_host.fused_stage(desc, [1.0], [1.0], [Q4], dt, math=math)
    U6 = emu_host.fused_stage(desc, [1.0], [1.0], [Q6], dt, math=math, ghosts=6)
    in4 = (slice(None),) + (slice(4, -4),) * desc.dim
    in6 = (slice(None),) + (slice(6, -6),) * desc.dim
    assert np.array_equal(U4[in4], U6[in6])
    if math == 0:
        Fo, So = oracle_lib.compute_flux_and_source(desc, Q4, dt)
        assert all(np.array_equal(F6[a], Fo[a]) for a in range(desc.dim))


@pytest.mark.parametrize("name", ["ss2d", "fe2d"])
def test_emulated_bulk_copy_staging_path(name, oracle_lib, monkeypatch):
    """The bulk-copy staging of the load phase (hb2_sweep.cuh: rows issued by warp 0 at the top of an iteration, consumed at
    its end; HB2_BULK_STAGE=1, off by default because it measured slower) produces the same bits as the per-thread cp.async
    path: fluxes and the fused stage against the oracle, long pencils with several segments included."""
    monkeypatch.setenv("HB2_BULK_STAGE", "1")
    desc, U = make_case(name, "random")
    Q = pb.pad_periodic(U)
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, 1.0e-3)
    Fe, Se = emu_host.flux_and_source(desc, Q, 1.0e-3, math=0)
    for a in range(desc.dim):
        assert np.array_equal(Fe[a], Fo[a])
    Uo = oracle_lib.advance_stage(desc, [1.0], [1.0], [Q], [Fo], [So])
    Ue = emu_host.fused_stage(desc, [1.0], [1.0], [Q], 1.0e-3, math=0, seg_len=8)
    assert np.array_equal(interior(desc, Ue), interior(desc, Uo))
