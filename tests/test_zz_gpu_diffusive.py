"""SURVEY row f4, GPU parity through the C ABI: the node-based sixth-order diffusive flux and the Navier-Stokes stage
update against the CPU oracle (oracle/oracle_diffusive.c) -- bit-identical (the diffusive unit is built with
-fmad=false and keeps the reference's operation order).  Written after the last GPU run of round 1: the same per-thread
functions pass tests/test_host_emu_diffusive.py on the CPU; this file is the parity test proper."""
import numpy as np
import pytest

from hamers_b200 import problems as pb
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TR = orc.Transport(mu=0.05, mu_v=0.02, c_p=3.5, c_v=2.5, Pr=0.72)


@pytest.fixture(autouse=True, params=[1, 0], ids=["march", "gridstride"])
def diff_kernel_form(request, monkeypatch):
    """Every test runs over both forms of the 3-D kernels: the marching ones with the asynchronous load pipeline
    (hb2_diffusive_march.cuh, the default) and the grid-stride ones; a plan reads HB2_DIFF_MARCH when it is created."""
    monkeypatch.setenv("HB2_DIFF_MARCH", str(request.param))
    return request.param


def _plan(desc):
    from hamers_b200 import abi

    return abi.DiffusivePlan(desc.dim, desc.n, desc.dx, desc.gamma[0], TR.c_v, TR.mu, TR.mu_v, TR.c_p, TR.Pr).use_torch_stream()


def _state(dim, N, seed=5):
    U, dx, gam = pb.random_state(dim, N, seed=seed, shock=True)
    return orc.PatchDesc(dim=dim, n=N, gamma=gam, dx=dx), U


@pytest.mark.parametrize("dim,N", [(2, (24, 17)), (3, (13, 10, 12)), (2, (1, 1)), (3, (7, 1, 3)), (3, (64, 48, 40)), (2, (300, 200)), (3, (70, 23, 19)), (3, (27, 3, 2)), (3, (40, 20, 100)), (3, (5, 9, 70))])
def test_diffusive_flux_device_and_host_entry_points(dim, N, product_lib):
    import torch

    desc, U = _state(dim, N)
    Q = pb.pad_periodic(U, orc.GD)
    dt = 1.0e-3
    Fo = orc.compute_diffusive_flux(desc, TR, Q, dt)
    plan = _plan(desc)
    Qd = torch.from_numpy(Q).cuda()
    Fd = [torch.full((desc.neq,) + desc.side_shape(a), float("nan"), dtype=torch.float64, device="cuda") for a in range(dim)]
    plan.compute_diffusive_flux(Qd, dt, Fd)
    torch.cuda.synchronize()
    for a in range(dim):
        assert np.array_equal(Fd[a].cpu().numpy(), Fo[a]), f"dir {a}"
    assert plan.launch_count in (1 + dim, 2 + dim)     # (primitives,) one-pass node fluxes, one face kernel per direction
    Fh = plan.compute_diffusive_flux_host(Q, dt)
    for a in range(dim):
        assert np.array_equal(Fh[a], Fo[a]), f"host entry point, dir {a}"
    plan.close()


@pytest.mark.parametrize("dim,N", [(2, (24, 17)), (3, (13, 10, 12)), (2, (1, 1)), (3, (7, 1, 3)), (3, (64, 48, 40)), (2, (300, 200)), (3, (27, 3, 2))])
def test_midpoint_reconstructor_flux(dim, N, product_lib):
    """hb2_diffusive_plan_set_reconstructor(HB2_DIFF_MIDPOINT_SIXTH_ORDER): DiffusiveFluxReconstructorMidpointSixthOrder
    through the C ABI against the oracle, bit for bit; the flux-free route refuses the midpoint reconstructor; switching back
    gives the node flux again."""
    import torch

    from hamers_b200 import abi

    desc, U = _state(dim, N)
    Q = pb.pad_periodic(U, orc.GD)
    dt = 1.0e-3
    Fo = orc.compute_diffusive_flux_midpoint(desc, TR, Q, dt)
    plan = _plan(desc).set_reconstructor(abi.DIFF_MIDPOINT_SIXTH_ORDER)
    Qd = torch.from_numpy(Q).cuda()
    Fd = [torch.full((desc.neq,) + desc.side_shape(a), float("nan"), dtype=torch.float64, device="cuda") for a in range(dim)]
    l0 = plan.launch_count
    plan.compute_diffusive_flux(Qd, dt, Fd)
    torch.cuda.synchronize()
    for a in range(dim):
        assert np.array_equal(Fd[a].cpu().numpy(), Fo[a]), f"dir {a}"
    assert plan.launch_count - l0 == 1 + 2 * dim
    Fh = plan.compute_diffusive_flux_host(Q, dt)
    for a in range(dim):
        assert np.array_equal(Fh[a], Fo[a]), f"host entry point, dir {a}"
    with pytest.raises(abi.HamersB200Error):
        plan.divergence_accumulate(Qd, dt, 6, 1.0, torch.zeros_like(Qd))
    with pytest.raises(abi.HamersB200Error):
        plan.set_reconstructor(7)
    plan.set_reconstructor(abi.DIFF_NODE_SIXTH_ORDER)
    plan.compute_diffusive_flux(Qd, dt, Fd)
    torch.cuda.synchronize()
    Fn = orc.compute_diffusive_flux(desc, TR, Q, dt)
    for a in range(dim):
        assert np.array_equal(Fd[a].cpu().numpy(), Fn[a]), f"node again, dir {a}"
    plan.close()


@pytest.mark.parametrize("dim,g", [(2, 6), (3, 6), (3, 4)])
def test_ns_stage_update(dim, g, product_lib):
    import torch

    rng = np.random.default_rng(11)
    N = (33, 20, 9)[:dim]
    desc = orc.PatchDesc(dim=dim, n=N, gamma=(1.4,), dx=(0.1, 0.2, 0.3)[:dim])
    neq = desc.neq
    shape = tuple(n + 2 * g for n in reversed(N))
    plan = _plan(desc)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()        # noqa: E731
    for alpha, beta in (([1.0], [1.0]), ([0.75, 0.25], [0.0, 0.25]), ([1.0 / 3.0, 0.0, 2.0 / 3.0], [0.0, 0.0, 2.0 / 3.0])):
        nc = len(alpha)
        U = [rng.standard_normal((neq,) + shape) for _ in range(nc)]
        Fc = [[rng.standard_normal((neq,) + desc.side_shape(a)) for a in range(dim)] for _ in range(nc)]
        Fd = [[rng.standard_normal((neq,) + desc.side_shape(a)) for a in range(dim)] for _ in range(nc)]
        S = [rng.standard_normal((neq,) + desc.cell_shape) for _ in range(nc)]
        Uo = orc.advance_stage_ns(desc, g, alpha, beta, U, Fc, Fd, S)
        out = torch.full((neq,) + shape, float("nan"), dtype=torch.float64, device="cuda")
        plan.advance_stage_ns(g, alpha, beta, [dev(u) for u in U], [[dev(f) for f in Fm] for Fm in Fc],
                              [[dev(f) for f in Fm] for Fm in Fd], [dev(s) for s in S], out)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        inner = (slice(None),) + (slice(g, -g),) * dim
        assert np.array_equal(got[inner], Uo[inner])
        got[inner] = np.nan
        assert np.isnan(got).all()
    plan.close()


def test_viscous_stage_from_both_reconstructors(product_lib):
    """One forward-Euler Navier-Stokes stage assembled like NavierStokes::computeFluxesAndSourcesOnPatch +
    advanceSingleStepOnPatch: convective flux from the WCNS5-JS / HLLC-HLL plan (ghost 4 view of the state), diffusive
    flux from the sixth-order plan (ghost 6), combined by hb2_advance_stage_ns_dev; against the oracle's same assembly."""
    import torch
    from hamers_b200 import abi

    N = (20, 16, 12)
    desc, U = _state(3, N)
    dt = 2.0e-4
    Q6, Q4 = pb.pad_periodic(U, 6), pb.pad_periodic(U, 4)
    Fc_o, S_o = orc.compute_flux_and_source(desc, Q4, dt)
    Fd_o = orc.compute_diffusive_flux(desc, TR, Q6, dt)
    Uo = orc.advance_stage_ns(desc, 6, [1.0], [1.0], [Q6], [Fc_o], [Fd_o], [S_o])
    cplan = abi.Plan(3, N, species_gamma=desc.gamma, dx=desc.dx, math=abi.MATH_EXACT).use_torch_stream()
    cplan6 = abi.Plan(3, N, species_gamma=desc.gamma, dx=desc.dx, math=abi.MATH_EXACT, num_ghosts=6).use_torch_stream()
    dplan = _plan(desc)
    Q6d, Q4d = torch.from_numpy(Q6).cuda(), torch.from_numpy(Q4).cuda()
    Fc = [torch.zeros((5,) + desc.side_shape(a), dtype=torch.float64, device="cuda") for a in range(3)]
    Fd = [torch.zeros((5,) + desc.side_shape(a), dtype=torch.float64, device="cuda") for a in range(3)]
    S = torch.zeros((5,) + desc.cell_shape, dtype=torch.float64, device="cuda")
    cplan.compute_flux_and_source(Q4d, dt, Fc, S)
    # the same reconstructor on the six-ghost array (plan with num_ghosts = 6) and on the extracted four-ghost view
    Fc6 = [torch.zeros_like(f) for f in Fc]
    S6 = torch.zeros_like(S)
    cplan6.compute_flux_and_source(Q6d, dt, Fc6, S6)
    view = torch.zeros_like(Q4d)
    dplan.extract_view(Q6d, 4, view)
    torch.cuda.synchronize()
    assert torch.equal(view, Q4d) and torch.equal(S6, S) and all(torch.equal(a, b) for a, b in zip(Fc6, Fc))
    dplan.compute_diffusive_flux(Q6d, dt, Fd)
    out = torch.zeros_like(Q6d)
    dplan.advance_stage_ns(6, [1.0], [1.0], [Q6d], [Fc], [Fd], [S], out)
    torch.cuda.synchronize()
    inner = (slice(None),) + (slice(6, -6),) * 3
    assert np.array_equal(out.cpu().numpy()[inner], Uo[inner])
    cplan.close()
    cplan6.close()
    dplan.close()


@pytest.mark.parametrize("dim,N", [(2, (24, 17)), (3, (13, 10, 12))])
def test_diffusive_reconstructor_class_matches_oracle(dim, N, tmp_path, product_lib):
    """DiffusiveFluxReconstructorNodeSixthOrder_B200 driven like NavierStokes::computeFluxesAndSourcesOnPatch does
    (tests/host_cpp/test_diffusive.cpp), host buffers, against the oracle."""
    from test_host_cpp import TRANSPORT, run_diffusive_driver

    U, dx, gam = pb.random_state(dim, N, seed=5, shock=True)
    Q = pb.pad_periodic(U, 6)
    dt = 7.5e-4
    r, fout = run_diffusive_driver(tmp_path, dim, N, Q, dx, gam[0], dt)
    assert r.returncode == 0, r.stdout + r.stderr
    desc = orc.PatchDesc(dim=dim, n=N, gamma=gam, dx=dx)
    t = TRANSPORT
    tr = orc.Transport(mu=t["mu"], mu_v=t["mu_v"], c_p=t["c_p"], c_v=1.0 / (gam[0] - 1.0) * t["R"], Pr=t["Pr"])
    Fo = orc.compute_diffusive_flux(desc, tr, Q, dt)
    out = np.fromfile(fout)
    pos = 0
    for a in range(dim):
        k = Fo[a].size
        assert np.array_equal(out[pos:pos + k].reshape(Fo[a].shape), Fo[a]), f"dir {a}"
        pos += k
    assert pos == out.size


def _oracle_ns_step(desc, tr, U, dt):
    """SSP-RK3 step of the Navier-Stokes patch strategy composed from the oracle's pieces (what NavierStokesLevel runs on
    the GPU): per stage flux + source of the newest state, diffusive flux, conservative update, periodic ghost fill."""
    from hamers_b200 import abi

    inner = (slice(None),) + (slice(6, -6),) * desc.dim
    states = [pb.pad_periodic(U, 6)]
    for s in range(3):
        m = s + 1
        newest = states[-1] if s < 2 else states[2]
        Fc, S = orc.compute_flux_and_source(desc, pb.pad_periodic(np.ascontiguousarray(newest[inner]), 4), dt)
        Fd = orc.compute_diffusive_flux(desc, tr, newest, dt)
        none = [None] * (m - 1)
        Uo = orc.advance_stage_ns(desc, 6, list(abi.SSPRK3_ALPHA[s][:m]), list(abi.SSPRK3_BETA[s][:m]), states[:m],
                                  none + [Fc], none + [Fd], none + [S])
        new = pb.pad_periodic(np.ascontiguousarray(Uo[inner]), 6)
        if s < 2:
            states.append(new)
        else:
            return new[inner]


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("dim,N", [(3, (16, 12, 10)), (2, (20, 14))])
def test_navier_stokes_level_steps_match_the_oracle_composition(dim, N, math, product_lib):
    """math = 0: materialised fluxes, reference association -> bit-identical.  math = 1: fused convective stage + diffusive
    divergence accumulated on top (no convective flux array) -> <= 1e-12 (fast-build criterion)."""
    import torch
    from common import assert_fast_parity
    from hamers_b200.ns_level import NavierStokesLevel

    rng = np.random.default_rng(9)
    # smooth-ish positive state (a viscous step of white noise with a Mach-3 slab would need a tiny dt to stay positive)
    ax = [(np.arange(n) + 0.5) / n for n in N]
    X = np.meshgrid(*reversed(ax), indexing="ij")[::-1]
    rho = 1.0 + 0.2 * np.sin(2 * np.pi * sum(X)) + 0.01 * rng.standard_normal(X[0].shape)
    vel = [0.4 * np.cos(2 * np.pi * X[a]) + 0.01 * rng.standard_normal(X[0].shape) for a in range(dim)]
    p = 1.0 + 0.1 * np.cos(2 * np.pi * X[0])
    U = np.stack([rho] + [rho * v for v in vel] + [p / 0.4 + 0.5 * rho * sum(v * v for v in vel)])
    lvl = NavierStokesLevel(dim, N, species_gamma=1.4, species_R=1.0, species_mu=TR.mu, species_mu_v=TR.mu_v,
                            species_c_p=TR.c_p, species_Pr=TR.Pr, domain=(0.0, 1.0), math=math)
    desc = orc.PatchDesc(dim=dim, n=N, gamma=(1.4,), dx=lvl.dx)
    tr = orc.Transport(mu=TR.mu, mu_v=TR.mu_v, c_p=TR.c_p, c_v=1.0 / (1.4 - 1.0) * 1.0, Pr=TR.Pr)
    lvl.interior().copy_(torch.from_numpy(U))
    dt = 2.0e-4
    want = U
    for _ in range(2):
        lvl.rk_step(dt)
        want = _oracle_ns_step(desc, tr, want, dt)
    torch.cuda.synchronize()
    got = lvl.S[lvl.cur].cpu().numpy()
    inner = (slice(None),) + (slice(6, -6),) * dim
    assert np.isfinite(got).all()
    if math == 0:
        assert np.array_equal(got[inner], want)
    else:
        assert_fast_parity(got[inner], want, "two SSP-RK3 steps")
    assert np.array_equal(got, pb.pad_periodic(np.ascontiguousarray(got[inner]), 6))        # ghosts valid after the step
    # viscosity acts: the result differs from the inviscid step, and mass is conserved to round-off
    assert abs(got[inner][0].sum() - U[0].sum()) < 1.0e-12 * U[0].size
    lvl.close()


@pytest.mark.parametrize("dim,N,g", [(2, (24, 17), 6), (3, (13, 10, 12), 6), (3, (40, 33, 20), 4), (3, (70, 23, 19), 6), (3, (33, 9, 4), 6), (3, (40, 20, 100), 6), (3, (5, 9, 70), 4)])
def test_flux_free_divergence_update(dim, N, g, product_lib):
    """hb2_diffusive_divergence_accumulate_dev == hb2_compute_diffusive_flux_dev + hb2_diffusive_accumulate_dev bit for bit,
    and both equal U + beta (-div F_d) formed from the oracle's side fluxes to round-off."""
    import torch

    desc, U = _state(dim, N)
    Q6 = pb.pad_periodic(U, orc.GD)
    rng = np.random.default_rng(2)
    base = rng.standard_normal((desc.neq,) + tuple(n + 2 * g for n in reversed(N)))
    dt, beta = 3.0e-4, 2.0 / 3.0
    plan = _plan(desc)
    Qd = torch.from_numpy(Q6).cuda()
    Fd = [torch.zeros((desc.neq,) + desc.side_shape(a), dtype=torch.float64, device="cuda") for a in range(dim)]
    two, one = torch.from_numpy(base).cuda(), torch.from_numpy(base).cuda()
    plan.compute_diffusive_flux(Qd, dt, Fd)
    plan.accumulate(g, beta, Fd, two)
    plan.divergence_accumulate(Qd, dt, g, beta, one)
    torch.cuda.synchronize()
    assert torch.equal(one, two)
    Fo = orc.compute_diffusive_flux(desc, TR, Q6, dt)
    inner = (slice(None),) + ((slice(g, -g),) * dim if g else ())
    want = base[inner].copy()
    for a in range(dim):
        want += beta * (-np.diff(Fo[a], axis=dim - a) / desc.dx[a])
    got = one.cpu().numpy()
    assert np.abs(got[inner] - want).max() <= 1.0e-13 * max(1.0, np.abs(want).max())
    outside = got.copy()
    outside[inner] = base[inner]
    assert np.array_equal(outside, base)                     # ghosts untouched
    plan.close()


@pytest.mark.parametrize("N,g", [((13, 10, 12), 6), ((40, 33, 20), 4), ((70, 23, 19), 6), ((40, 20, 100), 6), ((5, 9, 70), 6)])
def test_flux_free_divergence_fast_math(N, g, diff_kernel_form, product_lib):
    """hb2_diffusive_plan_set_math(HB2_MATH_FAST): the re-associated update agrees with the reference-order one to 1e-12 of
    the update's magnitude (plus the rounding of U + update); 3-D marching kernels only -- the grid-stride form keeps the
    exact arithmetic and must stay bit-identical."""
    import torch

    from hamers_b200 import abi

    desc, U = _state(3, N)
    Q6 = pb.pad_periodic(U, orc.GD)
    rng = np.random.default_rng(2)
    base = rng.standard_normal((desc.neq,) + tuple(n + 2 * g for n in reversed(N)))
    dt, beta = 3.0e-4, 2.0 / 3.0
    plan = _plan(desc)
    Qd = torch.from_numpy(Q6).cuda()
    exact, fast = torch.from_numpy(base).cuda(), torch.from_numpy(base).cuda()
    plan.divergence_accumulate(Qd, dt, g, beta, exact)
    plan.set_math(abi.MATH_FAST)
    plan.divergence_accumulate(Qd, dt, g, beta, fast)
    torch.cuda.synchronize()
    ex, fa = exact.cpu().numpy(), fast.cpu().numpy()
    if not diff_kernel_form:
        assert np.array_equal(ex, fa)
    upd = ex - base
    assert np.abs(upd[1:]).max() > 0.0 and np.array_equal(fa[0], base[0])
    for e in range(1, desc.neq):
        assert np.abs(fa[e] - ex[e]).max() <= 1.0e-12 * np.abs(upd[e]).max() + 2.0 * np.finfo(float).eps * np.abs(base[e]).max(), e
    inner = (slice(None),) + (slice(g, -g),) * 3
    outside = fa.copy()
    outside[inner] = base[inner]
    assert np.array_equal(outside, base)                     # ghosts untouched
    plan.close()


@pytest.mark.parametrize("push", ["1", "0"])
def test_two_gpu_navier_stokes_level_matches_single_box(push):
    """Box boundaries are invisible: two ranks with six-wide halos -- stored straight into the neighbour's state array
    (hb2_push_boxes_dev over CUDA IPC / NVLink, push = 1) or exchanged over NCCL (push = 0) -- against the one-box run
    (skipped on one GPU; the schedule is covered on CPU by tests/test_level_gloo.py::test_six_wide_halo_exchange_gloo and
    ::test_push_boxes_fill_every_ghost, the store kernel on one GPU by tests/test_zz_gpu_push_boxes.py)."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(root, "tests", "multi_gpu_ns_check.py"), "--size", "40"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, HB2_NS_PUSH=push))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "bit-identical = True" in r.stdout and f"push={push == '1'}" in r.stdout


@pytest.mark.parametrize("math", [100, 101])
def test_convective_class_on_a_six_ghost_state(math, oracle_lib, tmp_path):
    """ConvectiveFluxReconstructorWCNS5_JS_HLLC_HLL_B200 handed patch data with six ghost cells (what the Navier-Stokes
    application allocates): same fluxes, source and fused stage as on the four-ghost state (tests/host_cpp/
    test_reconstructor.cpp with the +100 flag)."""
    import struct
    import subprocess

    from common import assert_fast_parity, make_case
    from test_host_cpp import build_driver

    exe = build_driver()
    desc, U = make_case("ss3d", "random")
    Q4, Q6 = pb.pad_periodic(U, 4), pb.pad_periodic(U, 6)
    dt = 7.5e-4
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as fh:
        fh.write(struct.pack("7i", desc.dim, *desc.n, desc.model, desc.ns, math))
        fh.write(struct.pack("8d", *(list(desc.gamma) + [0.0] * 3 + list(desc.dx) + [dt])))
        fh.write(np.ascontiguousarray(Q6).tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = np.fromfile(fout)
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q4, dt)
    Uo = oracle_lib.advance_stage(desc, [1.0], [1.0], [Q4], [Fo], [So])
    pos = 0
    for a in range(3):
        k = Fo[a].size
        Fg = out[pos:pos + k].reshape(Fo[a].shape)
        pos += k
        if math % 10 == 0:
            assert np.array_equal(Fg, Fo[a]), f"dir {a}"
        else:
            assert_fast_parity(Fg, Fo[a], f"dir {a}")
    pos += So.size
    Ug = out[pos:].reshape(Q6.shape)
    in4, in6 = (slice(None),) + (slice(4, -4),) * 3, (slice(None),) + (slice(6, -6),) * 3
    if math % 10 == 0:
        assert np.array_equal(Ug[in6], Uo[in4])
    else:
        assert_fast_parity(Ug[in6], Uo[in4], "fused stage")


@pytest.mark.parametrize("dim,N", [(3, (16, 12, 10)), (2, (20, 14))])
def test_navier_stokes_stable_dt(dim, N, product_lib):
    """NavierStokes::computeSpectralRadiusesAndStableDtOnPatch: acoustic radii from the convective plan (on the six-ghost
    state), diffusive radius from the diffusive plan, against the oracle; also where viscosity limits the step."""
    import torch
    from hamers_b200.ns_level import NavierStokesLevel

    U, _, _ = pb.random_state(dim, N, seed=5, shock=True)
    for mu, length in ((0.05, 1.0), (0.05, 1.0e-3)):
        lvl = NavierStokesLevel(dim, N, species_gamma=1.4, species_R=1.0, species_mu=mu, species_mu_v=0.02, species_c_p=3.5,
                                species_Pr=0.72, domain=(0.0, length))
        lvl.interior().copy_(torch.from_numpy(U))
        lvl.fill_ghosts(lvl.S[lvl.cur])
        dt = lvl.stable_dt(1.0)
        desc = orc.PatchDesc(dim=dim, n=N, gamma=(1.4,), dx=lvl.dx)
        tr = orc.Transport(mu=mu, mu_v=0.02, c_p=3.5, c_v=1.0 / (1.4 - 1.0) * 1.0, Pr=0.72)
        radii, dt_o, sr_diff = orc.ns_spectral_radii_and_dt(desc, tr, lvl.c_p_eos, pb.pad_periodic(U, 6))
        assert lvl.spectral_radii[:dim].tolist() == radii and lvl.spectral_radii[4] == sr_diff
        assert abs(dt - dt_o) <= 1.0e-15 * dt_o
        if length < 1.0:
            assert sr_diff > lvl.spectral_radii[3]            # viscosity-limited
        lvl.close()
