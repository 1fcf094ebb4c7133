"""The stage loop of hamers_b200/ns_level.py (which states enter which stage with which coefficients, which buffer is
overwritten when, both routes) on the CPU: abi.Plan / abi.DiffusivePlan are replaced by stand-ins that run the host
emulation of the very kernels (tests/emu_host.py) on CPU tensors, and two SSP-RK3 steps are compared with the oracle's
composition of NavierStokes::computeFluxesAndSourcesOnPatch + advanceSingleStepOnPatch + periodic ghost fill.
The GPU counterpart is tests/test_zz_gpu_diffusive.py::test_navier_stokes_level_steps_match_the_oracle_composition."""
import numpy as np
import pytest

import emu_host
from common import assert_fast_parity
from hamers_b200 import abi
from hamers_b200 import problems as pb
from oracle import oracle as orc


class EmuConvectivePlan:
    def __init__(self, dim, n, flow_model=0, species_gamma=(1.4,), dx=(1.0, 1.0, 1.0), math=0, scheme=0, num_ghosts=4, **kw):
        self.desc = orc.PatchDesc(dim=dim, n=tuple(n), gamma=tuple(species_gamma), dx=tuple(dx), scheme=scheme)
        self.math, self.g, self.launch_count = math, num_ghosts, 0

    def use_torch_stream(self):
        return self

    def close(self):
        pass

    def compute_flux_and_source(self, Q, dt, flux, source=None):
        F, S = emu_host.flux_and_source(self.desc, Q.numpy(), dt, math=self.math, ghosts=self.g,
                                        source=source.numpy() if source is not None else None)
        for a, f in enumerate(F):
            flux[a].numpy()[...] = f

    def fused_stage(self, alpha, beta, U_int, dt, U_out, push=None):
        assert push is None
        out = emu_host.fused_stage(self.desc, list(alpha), list(beta), [u.numpy().copy() for u in U_int], dt, math=self.math,
                                   ghosts=self.g)
        inner = (slice(None),) + (slice(self.g, -self.g),) * self.desc.dim
        U_out.numpy()[inner] = out[inner]                      # the kernel leaves the ghosts of U_out alone


class EmuDiffusivePlan:
    def __init__(self, dim, n, dx, species_gamma, species_c_v, species_mu, species_mu_v, species_c_p, species_Pr, device=-1):
        self.desc = orc.PatchDesc(dim=dim, n=tuple(n), gamma=(species_gamma,), dx=tuple(dx))
        self.tr = orc.Transport(mu=species_mu, mu_v=species_mu_v, c_p=species_c_p, c_v=species_c_v, Pr=species_Pr)
        self.dim, self.n, self.neq, self.launch_count = dim, tuple(n), dim + 2, 0

    def set_math(self, math):            # the emulation keeps the reference-order arithmetic (inside the fast tolerance)
        return self

    def use_torch_stream(self):
        return self

    def close(self):
        pass

    def side_shape(self, d):
        s = list(self.n)
        s[d] += 1
        return tuple(reversed(s))

    def compute_diffusive_flux(self, Q, dt, flux):
        for a, f in enumerate(emu_host.diffusive_flux(self.desc, self.tr, Q.numpy(), dt)):
            flux[a].numpy()[...] = f

    def divergence_accumulate(self, Q, dt, g, beta, U):
        assert Q.data_ptr() != U.data_ptr()
        emu_host.diff_divergence_accumulate(self.desc, self.tr, Q.numpy(), dt, g, beta, U.numpy())

    def advance_stage_ns(self, g, alpha, beta, U_int, Fc_int, Fd_int, S_int, U_out):
        def np_rows(rows, nested):
            return [None if r is None else ([x.numpy() for x in r] if nested else r.numpy()) for r in rows]
        out = emu_host.advance_stage_ns(self.desc, self.tr, g, list(alpha), list(beta), [u.numpy().copy() for u in U_int],
                                        np_rows(Fc_int, True), np_rows(Fd_int, True), np_rows(S_int, False))
        inner = (slice(None),) + (slice(g, -g),) * self.dim
        U_out.numpy()[inner] = out[inner]

    def fill_ghosts_periodic(self, U, mask=7):
        emu_host.diff_fill_periodic(self.desc, self.tr, U.numpy(), mask)


def oracle_ns_step(desc, tr, U, dt):
    inner = (slice(None),) + (slice(6, -6),) * desc.dim
    states = [pb.pad_periodic(U, 6)]
    for s in range(3):
        m = s + 1
        newest = states[-1]
        Fc, S = orc.compute_flux_and_source(desc, pb.pad_periodic(np.ascontiguousarray(newest[inner]), 4), dt)
        Fd = orc.compute_diffusive_flux(desc, tr, newest, dt)
        none = [None] * (m - 1)
        Uo = orc.advance_stage_ns(desc, 6, list(abi.SSPRK3_ALPHA[s][:m]), list(abi.SSPRK3_BETA[s][:m]), states[:m],
                                  none + [Fc], none + [Fd], none + [S])
        states.append(pb.pad_periodic(np.ascontiguousarray(Uo[inner]), 6))
    return states[-1][inner]


@pytest.mark.parametrize("dim,N,math,scheme", [(3, (10, 8, 7), 0, 0), (3, (10, 8, 7), 1, 0), (2, (14, 9), 0, 0), (2, (14, 9), 1, 0),
                                                (3, (9, 8, 7), 0, 2)])          # last: WCNS6_LD, what the shipped TGV deck selects
def test_navier_stokes_level_loop_on_emulated_kernels(dim, N, math, scheme, monkeypatch):
    import torch
    from hamers_b200 import ns_level

    monkeypatch.setattr(abi, "Plan", EmuConvectivePlan)
    monkeypatch.setattr(abi, "DiffusivePlan", EmuDiffusivePlan)
    rng = np.random.default_rng(9)
    ax = [(np.arange(n) + 0.5) / n for n in N]
    X = np.meshgrid(*reversed(ax), indexing="ij")[::-1]
    rho = 1.0 + 0.2 * np.sin(2 * np.pi * sum(X)) + 0.01 * rng.standard_normal(X[0].shape)
    vel = [0.4 * np.cos(2 * np.pi * X[a]) + 0.01 * rng.standard_normal(X[0].shape) for a in range(dim)]
    p = 1.0 + 0.1 * np.cos(2 * np.pi * X[0])
    U = np.stack([rho] + [rho * v for v in vel] + [p / 0.4 + 0.5 * rho * sum(v * v for v in vel)])
    lvl = ns_level.NavierStokesLevel(dim, N, species_gamma=1.4, species_R=1.0, species_mu=0.05, species_mu_v=0.02,
                                     species_c_p=3.5, species_Pr=0.72, domain=(0.0, 1.0), math=math, device="cpu",
                                     scheme=scheme)
    desc = orc.PatchDesc(dim=dim, n=N, gamma=(1.4,), dx=lvl.dx, scheme=scheme)
    tr = orc.Transport(mu=0.05, mu_v=0.02, c_p=3.5, c_v=1.0 / (1.4 - 1.0) * 1.0, Pr=0.72)
    lvl.interior().copy_(torch.from_numpy(U))
    dt = 2.0e-4
    want = U
    for _ in range(2):
        lvl.rk_step(dt)
        want = oracle_ns_step(desc, tr, want, dt)
    got = lvl.S[lvl.cur].numpy()
    inner = (slice(None),) + (slice(6, -6),) * dim
    if math == 0:
        assert np.array_equal(got[inner], want)
    else:
        assert_fast_parity(got[inner], want, "two SSP-RK3 steps")
    assert np.array_equal(got, pb.pad_periodic(np.ascontiguousarray(got[inner]), 6))
    assert abs(got[inner][0].sum() - U[0].sum()) < 1.0e-12 * U[0].size
    assert np.abs(got[inner] - U).max() > 1.0e-5            # something happened


# ---- the same loop over SEVERAL ranks (gloo, CPU): box decomposition, six-wide one-shot exchange through the plan's multi-box
# ---- pack / unpack, local periodic fill (ghost-inclusive in the exchanged directions) by the emulated fill kernel ----------
def _region(lo, hi, dim, g=6):
    return (slice(None),) + tuple(slice(lo[a] + g, hi[a] + g) for a in reversed(range(dim)))


def _emu_box_table(self, boxes, offsets):
    return list(boxes), list(offsets)


def _emu_pack_boxes(self, U, table, buffer):
    for (lo, hi), off in zip(*table):
        part = U.numpy()[_region(lo, hi, self.desc.dim, self.g)]
        buffer.numpy()[off:off + part.size] = part.reshape(-1)         # component-major, x fastest: k_multibox's order


def _emu_unpack_boxes(self, U, table, buffer):
    for (lo, hi), off in zip(*table):
        view = U.numpy()[_region(lo, hi, self.desc.dim, self.g)]
        view[...] = buffer.numpy()[off:off + view.size].reshape(view.shape)


def _level_state(dim, N):
    rng = np.random.default_rng(9)
    ax = [(np.arange(n) + 0.5) / n for n in N]
    X = np.meshgrid(*reversed(ax), indexing="ij")[::-1]
    rho = 1.0 + 0.2 * np.sin(2 * np.pi * sum(X)) + 0.01 * rng.standard_normal(X[0].shape)
    vel = [0.4 * np.cos(2 * np.pi * X[a]) + 0.01 * rng.standard_normal(X[0].shape) for a in range(dim)]
    p = 1.0 + 0.1 * np.cos(2 * np.pi * X[0])
    return np.stack([rho] + [rho * v for v in vel] + [p / 0.4 + 0.5 * rho * sum(v * v for v in vel)])


def _ns_worker(rank, world, port, dim, N, math, steps, q):
    import os

    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hamers_b200 import ns_level

        EmuConvectivePlan.box_table = _emu_box_table
        EmuConvectivePlan.pack_boxes = _emu_pack_boxes
        EmuConvectivePlan.unpack_boxes = _emu_unpack_boxes
        abi.Plan, abi.DiffusivePlan = EmuConvectivePlan, EmuDiffusivePlan
        U = _level_state(dim, N)
        lvl = ns_level.NavierStokesLevel(dim, N, species_gamma=1.4, species_R=1.0, species_mu=0.05, species_mu_v=0.02,
                                         species_c_p=3.5, species_Pr=0.72, domain=(0.0, 1.0), math=math, device="cpu")
        assert lvl.dist is not None and not lvl.push
        d = lvl.decomp
        box = (slice(None),) + tuple(slice(d.lo[a], d.lo[a] + d.n[a]) for a in reversed(range(dim)))
        lvl.interior().copy_(torch.from_numpy(np.ascontiguousarray(U[box])))
        for _ in range(steps):
            lvl.rk_step(2.0e-4)
        got = lvl.S[lvl.cur].numpy()
        inner = (slice(None),) + (slice(6, -6),) * dim
        q.put((rank, tuple(d.lo), tuple(d.n), got[inner].copy(), bool(np.isfinite(got).all())))
    except Exception as e:      # noqa: BLE001 -- reported to the parent, which fails the test
        q.put((rank, None, None, repr(e), False))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim,N,world,math", [(3, (7, 8, 12), 2, 0), (2, (14, 12), 2, 0), (3, (7, 12, 12), 4, 0),
                                               (3, (7, 8, 12), 2, 1)])
def test_navier_stokes_level_over_several_ranks_equals_the_one_box_oracle(dim, N, world, math):
    """World-size 2 / 4 (gloo): every rank advances its box of the periodic level with the emulated kernels; the gathered
    boxes equal the oracle's composition on the WHOLE level bit for bit in the reference-order route (a box boundary is
    invisible: what fill_schedule->fillData guarantees, RungeKuttaLevelIntegrator.cpp:1568/1701), including the edge and
    corner ghosts that a local periodic fill has to produce from just-received cells."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    steps = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ns_worker, args=(r, world, port, dim, N, math, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    U = _level_state(dim, N)
    got = np.full_like(U, np.nan)
    for rank, lo, n, part, finite in res:
        assert lo is not None, f"rank {rank}: {part}"
        assert finite, f"rank {rank}: non-finite cells in the ghost box (ghosts never filled)"
        got[(slice(None),) + tuple(slice(lo[a], lo[a] + n[a]) for a in reversed(range(dim)))] = part
    dx = tuple(1.0 / n for n in N)
    desc = orc.PatchDesc(dim=dim, n=N, gamma=(1.4,), dx=dx)
    tr = orc.Transport(mu=0.05, mu_v=0.02, c_p=3.5, c_v=1.0 / (1.4 - 1.0) * 1.0, Pr=0.72)
    want = U
    for _ in range(steps):
        want = oracle_ns_step(desc, tr, want, 2.0e-4)
    if math == 0:
        assert np.array_equal(got, want)
    else:
        assert_fast_parity(got, want, "two SSP-RK3 steps over several ranks")
