#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native WCNS5-JS / HLLC-HLL path.

Metric (BASELINE.json): FP64 cell-updates/s per RK3 stage, WCNS5-JS 3D Euler, single-species,
512^3 periodic uniform level (tests/3D_convergence_test_single_species scaled up, SURVEY.md 8d "M1").
One cell-update = one interior cell advanced through one RK stage (flux in all directions + update).
A "step" here is one SSP-RK3 time step = three passes of the hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--math fast|exact]
                  [--scaling strong|weak] [--model ss|fe] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the level is cut into one box per rank and
the width-4 halos are exchanged over NVLink every stage.  `--impl reference` times the CPU oracle
(a restatement of the reference's algorithm; the real reference needs SAMRAI+MPI+HDF5 and cannot be
built here) on the host cores, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per cell-update, single-species 3D / five-eqn 3D
ALGO = {
    "ss": {"flops": 3.8e3, "bytes": 106.7, "flops_sweep": 1.217e3, "flops_sensor": 73.0, "flops_rk": 75.0},
    "fe": {"flops": 5.6e3, "bytes": 165.0, "flops_sweep": 1.80e3, "flops_sensor": 90.0, "flops_rk": 110.0},
}


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_ic_device(level, model):
    """Convergence-test IC (ConvergenceSingleSpecies.cpp:145-174 / ConvergenceFiveEqnAllaire.cpp:176-219) of this
    rank's box, generated on the device (synthetic data)."""
    import torch

    xs = [torch.as_tensor(c, dtype=torch.float64, device="cuda") for c in level.local_coordinates()]
    s = (xs[0][None, None, :] + xs[1][None, :, None]) + xs[2][:, None, None]
    sin = torch.sin(np.pi * s)
    dim = 3
    if model == "ss":
        rho = 1.0 + 0.5 * sin
        E = 1.0 / (7.0 / 5.0 - 1.0) + 0.5 * rho * float(dim)
        comps = [rho, rho, rho, rho, E]
    else:
        Z1 = 0.5 + 0.25 * sin
        Z2 = 1.0 - Z1
        Zr1, Zr2 = Z1 * 2.0, Z2 * 1.0
        rho = Zr1 + Zr2
        gm = 1.0 / (Z1 / (8.0 / 5.0 - 1.0) + Z2 / (7.0 / 5.0 - 1.0)) + 1.0
        E = 1.0 / (gm - 1.0) + 0.5 * rho * float(dim)
        comps = [Zr1, Zr2, rho, rho, rho, E, Z1, Z2]
    inter = level.interior()
    for c, v in enumerate(comps):
        inter[c].copy_(v)
    del sin, s


def bit_checksum(t):
    """Order-independent checksum of a float64 tensor: the sum of the 64-bit patterns modulo 2^64.  Two runs whose cells
    hold the same bits give the same number whatever the decomposition into boxes."""
    import torch

    return int(t.contiguous().view(torch.int64).sum().item())


def level_parity(level, model, dist, n_total_steps, dt):
    """What xfer::RefineSchedule::fillData guarantees (RungeKuttaLevelIntegrator.cpp:1568, 1701): a box boundary is
    invisible.  (i) the error norms of the reference's own acceptance statistic
    (problems/Euler/error_statistics/ConvergenceSingleSpecies.cpp:199-249: L1, L2, Linf of rho against the advected
    exact solution; five-eqn: of Z_1, ConvergenceFiveEqnAllaire.cpp) after all steps run so far, reduced over ranks like
    the reference's MPI_SUM / MPI_MAX; (ii) a bit checksum of the whole state (sum of the 64-bit patterns mod 2^64, summed
    over ranks)."""
    import torch

    t = n_total_steps * dt
    xs = [torch.as_tensor(c, dtype=torch.float64, device="cuda") for c in level.local_coordinates()]
    s = (xs[0][None, None, :] + xs[1][None, :, None]) + xs[2][:, None, None]
    inter = level.S[level.cur][level._interior_slices()]
    if model == "ss":
        exact = 1.0 + 0.5 * torch.sin(np.pi * (s - 3.0 * t))
        num = inter[0]
    else:
        exact = 0.5 + 0.25 * torch.sin(np.pi * (s - 3.0 * t))
        num = inter[-2]
    err = (exact - num).abs()
    dvol = float(np.prod(level.dx))
    acc = torch.stack([err.sum() * dvol, (err * err).sum() * dvol, torch.tensor(dvol * err.numel(), dtype=torch.float64, device="cuda")])
    emax = err.max().reshape(1)
    del err, exact, s
    cks = torch.tensor([bit_checksum(inter)], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        dist.all_reduce(emax, op=dist.ReduceOp.MAX)
        dist.all_reduce(cks, op=dist.ReduceOp.SUM)          # int64 addition wraps: still the sum modulo 2^64
    return {"time": t, "steps_total": n_total_steps, "L1_error": float(acc[0] / acc[2]), "L2_error": float(torch.sqrt(acc[1] / acc[2])),
            "Linf_error": float(emax[0]), "checksum": int(cks[0]) & 0xFFFFFFFFFFFFFFFF,
            "statistic": "rho vs exact advected solution (ConvergenceSingleSpecies.cpp:199-249)" if model == "ss"
            else "Z_1 vs exact advected solution (ConvergenceFiveEqnAllaire.cpp)"}


def single_box_replica(args, Nglob, flow_model, gam, math, scheme, n_total_steps, dt, model):
    """The same level as ONE box on this GPU (the N = 1 configuration: periodic-fill kernel, no exchange), advanced by the
    same number of steps: its checksum and error norms are what a multi-GPU run must reproduce."""
    import torch

    from hamers_b200.level import UniformLevel

    one = UniformLevel(3, Nglob, flow_model=flow_model, species_gamma=gam, math=math, scheme=scheme, distributed=False)
    make_ic_device(one, model)
    for _ in range(n_total_steps):
        one.rk_step(dt)
    torch.cuda.synchronize()
    out = level_parity(one, model, None, n_total_steps, dt)
    one.close()
    del one
    torch.cuda.empty_cache()
    return out


def secondary_measurements(args, torch, dist, rank, world, local_rank):
    """Numbers next to the headline, same timing rules (CUDA events, 3 warm-up steps, max over ranks):
      ns     BASELINE.json config 5: 3-D single-species Navier-Stokes Taylor-Green vortex (Re 1600, M 0.1), box-partitioned
             over the ranks, WCNS5-JS + SIXTH_ORDER diffusive flux, fast route (N = 1: also WCNS6_LD, the shipped deck's choice)
      shock  N = 1: the branch-coverage state M2 of SURVEY.md 8d at 256^3 (random state + Mach-3 slab: sensor fires, HLL
             blend and first-order fallback are executed) -- one fused stage, repeated on the same state
      fe     N = 1: config 3 at 384^3 (five-equation Allaire model)"""
    from hamers_b200 import abi
    from hamers_b200.level import PROCESS_GRIDS
    from hamers_b200.ns_level import NavierStokesLevel

    out = {}

    def timed(fn, steps, warm=3):
        for _ in range(warm):
            fn()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / steps

    # ---- Navier-Stokes Taylor-Green vortex ---------------------------------------------------------------------------------
    n = args.ns_size
    grid = PROCESS_GRIDS[3][world]
    for scheme_name, scheme in (("WCNS5_JS", abi.WCNS5_JS),) + ((("WCNS6_LD", abi.WCNS6_LD),) if world == 1 else ()):
        lvl = NavierStokesLevel(3, (n, n, n), species_gamma=1.4, species_R=1.0, species_mu=1.0 / 1600.0, species_mu_v=0.0,
                                species_c_p=3.5, species_Pr=0.71, domain=(0.0, 2.0 * np.pi), math=abi.MATH_FAST, scheme=scheme)
        x, y, z = [torch.as_tensor(c, dtype=torch.float64, device="cuda") for c in lvl.coordinates()]
        X, Y, Z = x[None, None, :], y[None, :, None], z[:, None, None]
        M0 = 0.1
        u = torch.sin(X) * torch.cos(Y) * torch.cos(Z)
        v = -torch.cos(X) * torch.sin(Y) * torch.cos(Z)
        p = 1.0 / (1.4 * M0 * M0) + (torch.cos(2 * X) + torch.cos(2 * Y)) * (torch.cos(2 * Z) + 2.0) / 16.0
        rho = torch.ones_like(u)
        inter = lvl.interior()
        for c, f in enumerate([rho, rho * u, rho * v, torch.zeros_like(u), p / 0.4 + 0.5 * rho * (u * u + v * v)]):
            inter[c].copy_(f.expand_as(inter[c]))
        del u, v, p, rho
        dt = 0.2 * lvl.dx[0] / (1.0 + 1.0 / M0)
        l0 = lvl.launch_count
        ms = timed(lambda: lvl.rk_step(dt), args.ns_steps)
        state = lvl.S[lvl.cur][lvl._interior_slices()]
        ke = (0.5 * (state[1] ** 2 + state[2] ** 2 + state[3] ** 2) / state[0]).sum().reshape(1)
        cks = torch.tensor([bit_checksum(state)], dtype=torch.int64, device="cuda")
        if dist is not None:
            dist.all_reduce(ke, op=dist.ReduceOp.SUM)
            dist.all_reduce(cks, op=dist.ReduceOp.SUM)
        key = "ns" if scheme == abi.WCNS5_JS else "ns_wcns6ld"
        out[key] = {"workload": f"3D single-species Navier-Stokes, Taylor-Green vortex Re=1600 M=0.1, periodic {n}^3 over "
                                f"{world} box(es) {list(grid)}, {scheme_name}_HLLC_HLL + SIXTH_ORDER, SSP-RK3, fast route",
                    "value": float(n) ** 3 * 3 / (ms * 1e-3), "unit": "cell-updates/s", "ms_per_step": ms, "steps": args.ns_steps,
                    "gpu_launches_per_step": (lvl.launch_count - l0) / (args.ns_steps + 3),
                    "mean_kinetic_energy": float(ke[0]) / float(n) ** 3, "checksum": int(cks[0]) & 0xFFFFFFFFFFFFFFFF,
                    "finite": bool(torch.isfinite(state).all()),
                    "ghost_fill": ("periodic fill kernel" if world == 1 else
                                   "hb2_push_boxes_dev: six-wide halos stored into the neighbours' arrays over CUDA IPC / NVLink"
                                   if lvl.push else "pack -> NCCL send/recv -> unpack")}
        lvl.close()
        del lvl, state
        torch.cuda.empty_cache()
    if world > 1:
        return out

    # ---- shocked field -------------------------------------------------------------------------------------------------------
    from hamers_b200 import problems as pb

    for model, name in ((abi.SINGLE_SPECIES, "shock"),):
        n = args.shock_size
        U, dx, gam = pb.random_state(3, (n, n, n), model=model, seed=20261017, shock=True)
        plan = abi.Plan(3, (n, n, n), flow_model=model, species_gamma=gam, dx=dx, math=abi.MATH_FAST).use_torch_stream()
        Q = torch.from_numpy(pb.pad_periodic(U)).cuda()
        out_t = torch.zeros_like(Q)
        dtq = 1.0e-3 * dx[0]
        ms = timed(lambda: plan.fused_stage([1.0], [1.0], [Q], dtq, out_t), 10)
        # how many faces take the data-dependent branches: the sensor decisions of the plan (one byte per cell, bit d = HLLC-HLL
        # on the low face in direction d) are not exported; count them with the oracle-free criterion instead: faces whose
        # fused result differs from a run with the HLL blend disabled cannot be had here, so report the state's statistics
        out[name] = {"workload": f"3D single-species Euler, WCNS5_JS_HLLC_HLL, one fused stage on the random + Mach-3-slab state "
                                 f"(SURVEY 8d M2) at {n}^3, fast build", "value": float(n) ** 3 / (ms * 1e-3),
                     "unit": "cell-updates/s", "ms_per_stage": ms, "finite": bool(torch.isfinite(out_t[:, 4:-4, 4:-4, 4:-4]).all())}
        # the same size on the smooth field, for the ratio
        Us, dxs, gs = pb.convergence_single_species(3, n)
        Q.copy_(torch.from_numpy(pb.pad_periodic(Us)))
        ms_s = timed(lambda: plan.fused_stage([1.0], [1.0], [Q], dtq, out_t), 10)
        out[name]["smooth_same_size_value"] = float(n) ** 3 / (ms_s * 1e-3)
        out[name]["shock_over_smooth"] = ms_s / ms
        plan.close()
        del Q, out_t
        torch.cuda.empty_cache()

    # ---- five-equation model ---------------------------------------------------------------------------------------------------
    from hamers_b200.level import UniformLevel

    n = args.fe_size
    lv = UniformLevel(3, (n, n, n), flow_model=abi.FIVE_EQN_ALLAIRE, species_gamma=(8.0 / 5.0, 7.0 / 5.0), math=abi.MATH_FAST,
                      distributed=False)
    make_ic_device(lv, "fe")
    dtf = 0.001 * lv.dx[0]
    ms = timed(lambda: lv.rk_step(dtf), 5)
    out["fe"] = {"workload": f"3D five-equation Allaire, WCNS5_JS_HLLC_HLL, SSP-RK3, periodic {n}^3 (config 3 scaled up), fast build",
                 "value": float(n) ** 3 * 3 / (ms * 1e-3), "unit": "cell-updates/s", "ms_per_step": ms,
                 "fp64_frac_of_5.6kFLOP_roofline": None, "finite": bool(torch.isfinite(lv.interior()).all())}
    lv.close()
    del lv
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    """CPU arm: the oracle (reference-structured restatement, gcc -O3, one OpenMP thread per core standing in
    for one MPI rank, 32^3 patches) on a bounded sample of the same workload."""
    rank, _, world = env_rank()
    if rank != 0:
        return
    from hamers_b200 import problems as pb
    from oracle import oracle as orc

    orc.build()
    cores = os.cpu_count() or 1
    N = args.ref_size
    model = 0 if args.model == "ss" else 1
    if model == 0:
        U, dx, gam = pb.convergence_single_species(3, N)
    else:
        U, dx, gam = pb.convergence_five_eqn(3, N)
    lvl = orc.PatchDesc(dim=3, n=(N,) * 3, model=model, ns=len(gam), gamma=gam, dx=dx)
    dt = 0.001 * dx[0]
    patch = (min(32, N),) * 3
    for _ in range(args.warmup if args.warmup > 0 else 0):
        orc.level_advance(lvl, patch, U, dt, 1, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.level_advance(lvl, patch, U, dt, 1, nthreads=cores)
    el = time.perf_counter() - t0
    value = N ** 3 * 3 * args.steps / el
    sample = f"{N}^3 box in {patch[0]}^3 patches, {args.steps} SSP-RK3 steps ({3 * args.steps} stages), {cores} OpenMP threads"
    line = {
        "impl": "reference", "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, sample_override=sample),
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sample_override=None):
    sch = {"js": "WCNS5_JS_HLLC_HLL", "z": "WCNS5_Z_HLLC_HLL", "ld": "WCNS6_LD_HLLC_HLL"}[getattr(args, "scheme", "js")]
    name = ("3D single-species Euler" if args.model == "ss" else "3D five-equation Allaire") + \
           f", {sch}, SSP-RK3, periodic uniform level {args.size}^3 (convergence-test IC scaled up)"
    cfg = {"workload": name, "size": args.size, "math": args.math, "dt": "0.001*dx", "ghosts": 4,
           "l2_policy": "inputs (>= 5 GB per state) exceed the 126 MB L2; no flush needed",
           "step": "one SSP-RK3 time step = 3 hot-path passes (ghost fill + sensor + x/y/z sweeps with fused RK update)"}
    if sample_override:
        cfg["sample"] = sample_override
    return cfg


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity), so that the pinned host boxes of the
    e2e leg are first-touched on the GPU's own NUMA node (torchrun does not bind ranks).  Returns the previous affinity (to be
    restored for the CPU baseline, which uses every core) or None when NVML cannot say."""
    if os.environ.get("HB2_BENCH_AFFINITY", "1") == "0":
        return None
    try:
        import pynvml

        before = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        return before
    except Exception:
        return None


def run_ours(args):
    import torch

    rank, local_rank, world = env_rank()
    affinity_before = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    from hamers_b200 import abi
    from hamers_b200.level import PROCESS_GRIDS, UniformLevel

    flow_model = abi.SINGLE_SPECIES if args.model == "ss" else abi.FIVE_EQN_ALLAIRE
    gam = (7.0 / 5.0,) if args.model == "ss" else (8.0 / 5.0, 7.0 / 5.0)
    math = abi.MATH_FAST if args.math == "fast" else abi.MATH_EXACT
    grid = PROCESS_GRIDS[3][world]
    if args.scaling == "strong":
        Nglob = (args.size,) * 3
    else:
        Nglob = tuple(args.size * g for g in grid)
    domain_len = 2.0
    scheme = {"js": abi.WCNS5_JS, "z": abi.WCNS5_Z, "ld": abi.WCNS6_LD}[args.scheme]
    level = UniformLevel(3, Nglob, flow_model=flow_model, species_gamma=gam, math=math, scheme=scheme)
    # keep dx = 2/size in weak scaling too (same physics per cell)
    make_ic_device(level, args.model)
    dx = level.dx[0]
    dt = 0.001 * dx
    ncell_local = int(np.prod(level.decomp.n))
    ncell_global = int(np.prod(Nglob))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        level.rk_step(dt)
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = level.plan.launch_count
    level.plan.set_profiling(True)
    level.plan.get_profile(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        level.rk_step(dt)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    prof = level.plan.get_profile(reset=True)
    level.plan.set_profiling(False)
    launches = level.plan.launch_count - launches0
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
        launches = int(tsum[1])
    sec = ms * 1e-3
    value = ncell_global * 3 * args.steps / sec

    # parity of the run itself: error norms against the exact solution + bit checksum; for N > 1 rank 0 then repeats the run
    # as one box (N = 1 configuration) and the two must agree bit for bit
    n_total = max(args.warmup, 0) + args.steps
    parity = level_parity(level, args.model, dist, n_total, dt)
    if world > 1 and not args.no_replica:
        if ncell_global <= 640 ** 3:
            rep = None
            if rank == 0:
                rep = single_box_replica(args, Nglob, flow_model, gam, math, scheme, n_total, dt, args.model)
            if rank == 0:
                parity["n1_checksum"] = rep["checksum"]
                parity["n1_L1_error"] = rep["L1_error"]
                parity["n1_L2_error"] = rep["L2_error"]
                parity["matches_n1"] = bool(rep["checksum"] == parity["checksum"])
                parity["n1_source"] = "same level advanced as ONE box on rank 0's GPU in this run (outside the timed region)"
            dist.barrier()
        else:
            parity["matches_n1"] = None
            parity["n1_source"] = "level too large for one GPU next to this rank's box: not replicated"
    elif world == 1:
        parity["matches_n1"] = True
        parity["n1_source"] = "this is the N = 1 run"

    # sanity of the run itself: the solution must stay finite and close to the advected wave
    rho_min = float(level.interior()[0].min())
    rho_max = float(level.interior()[0].max())
    finite = bool(torch.isfinite(level.interior()).all())

    line = None
    if rank == 0:
        algo = ALGO[args.model]
        # FP64 and HBM peaks
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        fp64_peak = abi.probe_fp64_peak(local_rank, 1.0) / 1e12
        # dominant kernel = the slowest sweep
        sweeps = {k: prof[k] for k in ("xsweep", "ysweep", "zsweep") if prof[k][1] > 0}
        dom = max(sweeps, key=lambda k: sweeps[k][0] / sweeps[k][1])
        dom_ms = sweeps[dom][0] / sweeps[dom][1]
        dom_flops = (algo["flops_sweep"] + (algo["flops_rk"] if dom == "zsweep" else 0.0)) * ncell_local
        kern = {k: {"avg_ms": v[0] / v[1], "launches": v[1], "share": v[0] / max(1e-30, sum(x[0] for x in prof.values()))}
                for k, v in prof.items() if v[1] > 0}
        per_rank_rate = value / world
        roofline = {
            "bound": "fp64", "kernel": dom, "achieved": dom_flops / (dom_ms * 1e-3) / 1e12, "peak": fp64_peak,
            "unit": "TFLOP/s", "traffic": None,
            "peak_source": "in-run DFMA probe (hb2_probe_fp64_peak); MEASURED_PEAKS.json has no FP64 figure; "
                           "bound is the FP64 pipe (AI ~36 FLOP/B >> ridge), 'hbm' view given beside it",
            "algorithmic_flops_per_launch": dom_flops,
            "stage": {"fp64_achieved_tflops": algo["flops"] * per_rank_rate / 1e12,
                      "fp64_frac": algo["flops"] * per_rank_rate / 1e12 / fp64_peak,
                      "hbm_achieved_gbs": algo["bytes"] * per_rank_rate / 1e9, "hbm_peak_gbs": hbm_peak,
                      "hbm_frac": algo["bytes"] * per_rank_rate / 1e9 / hbm_peak, "hbm_peak_source": hbm_src},
            "kernels": kern,
        }
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
        # DRAM traffic of the dominant kernel per launch: NOT measured in this run (ncu is never attached to a bench run); the
        # value is read from the committed ncu --set full capture of the same workload and labelled as such
        design_bytes = {"xsweep": 80.0, "ysweep": 120.0, "zsweep": 160.0}     # bytes per cell the three-sweep DESIGN moves
        try:
            with open(os.path.join(ROOT, "profiles", "r02_as_traffic.json")) as fh:
                tr = json.load(fh)
            if tr["size"] == args.size and tr["model"] == args.model and tr["math"] == args.math and world == 1:
                roofline["traffic"] = tr["dram_bytes_per_launch"][dom]
                roofline["traffic_unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)"
                roofline["traffic_source"] = "from committed capture, not measured in this run: " + tr["source"]
        except Exception:
            pass
        roofline["design_bytes_per_launch"] = design_bytes[dom] * ncell_local
        # algorithmic traffic of a whole STAGE (fused, no flux materialisation) is 80 / 120 / 120 B per cell for
        # stages 0 / 1 / 2; the three-sweep design adds the R round trips between the sweeps and two more reads of the state
        roofline["algorithmic_bytes_per_stage"] = {"stage0": 80.0 * ncell_local, "stage1": 120.0 * ncell_local, "stage2": 120.0 * ncell_local}
        # the same dominant kernel against the HBM roof (the secondary roof of this FP64-bound path): bytes the DESIGN moves in
        # that launch (x 80, y 120, z 160 B/cell with the R round trips) / its time
        dom_bytes = design_bytes[dom] * ncell_local
        roofline_hbm = {"bound": "hbm", "kernel": dom, "achieved": dom_bytes / (dom_ms * 1e-3) / 1e9, "peak": hbm_peak,
                        "unit": "GB/s", "peak_source": hbm_src, "traffic": roofline.get("traffic"),
                        "bytes": "design bytes of the launch (not SURVEY 8d's algorithmic stage bytes, given in roofline.algorithmic_bytes_per_stage)"}
        roofline_hbm["frac"] = roofline_hbm["achieved"] / roofline_hbm["peak"]
        line = {
            "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args), global_cells=list(Nglob), process_grid=list(grid),
                           cells_per_gpu=ncell_local, parallelism=f"box{world}"),
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "roofline_hbm": roofline_hbm,
            "sanity": {"finite": finite, "rho_min": rho_min, "rho_max": rho_max}, "parity": parity,
        }

    # ---- secondary measurements (config 5, shocked field, config 3) -----------------------------------
    if not args.no_secondary:
        sec = secondary_measurements(args, torch, dist, rank, world, local_rank)
        if rank == 0:
            if "fe" in sec:
                sec["fe"]["fp64_frac_of_5.6kFLOP_roofline"] = ALGO["fe"]["flops"] * sec["fe"]["value"] / 1e12 / line["roofline"]["peak"]
            line["secondary"] = sec

    # ---- end-to-end: host buffers through the C ABI (N = 1: advanceLevel on host memory) ----------
    if world == 1 and not args.no_e2e:
        n = args.e2e_size
        plan = abi.Plan(3, (n,) * 3, flow_model=flow_model, species_gamma=gam, dx=(domain_len / n,) * 3, math=math, scheme=scheme)
        ncomp = plan.ncomp
        host = torch.empty((ncomp,) + plan.ghost_shape, dtype=torch.float64).pin_memory()
        lvl2 = level if n == args.size else None
        if lvl2 is not None:
            host.copy_(level.S[level.cur])
        else:
            from hamers_b200 import problems as pb

            U, _, _ = (pb.convergence_single_species(3, n) if args.model == "ss" else pb.convergence_five_eqn(3, n))
            host.copy_(torch.from_numpy(pb.pad_periodic(U)))
        level.close()
        del level
        torch.cuda.empty_cache()
        hnp = host.numpy()
        dte = 0.001 * domain_len / n
        # slab thicknesses along z.  0 (default): 14 slabs of n/16 planes between two pairs of n/32 (the download of a slab
        # trails its upload by two slabs plus three stages, so thin slabs start the download stream early; the slabs around
        # the periodic seam finish after the upload has ended -- their download is the tail of the step -- and are thinner
        # still); K > 1: K equal slabs; 1: one patch through hb2_advance_level_host
        if args.e2e_patches == 0 and n % 32 == 0 and n // 32 >= 12:
            thin = n // 32
            sizes = [thin, thin] + [2 * thin] * 14 + [thin, thin]
        elif args.e2e_patches > 1 and n % args.e2e_patches == 0 and n // args.e2e_patches >= 16:
            sizes = [n // args.e2e_patches] * args.e2e_patches
        elif args.e2e_patches == 0 and n % 8 == 0 and n // 8 >= 16:
            sizes = [n // 8] * 8
        else:
            sizes = [n]
        npz = len(sizes)
        if npz == 1:
            for _ in range(1):
                plan.advance_level_host(hnp, dte)
            k0 = plan.launch_count
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                plan.advance_level_host(hnp, dte)
            el = time.perf_counter() - t0
            nbytes = host.numel() * 8
            launches = plan.launch_count - k0
            api = "hb2_advance_level_host (advanceLevel on pinned host memory: H2D U^n, 3 stages on device, D2H U^{n+1})"
            plan.close()
        else:
            # the host box as npz slabs along z (what a SAMRAI level of several patches per rank looks like): the uploads of the
            # later slabs overlap the first stage of the earlier ones, the downloads overlap the last stage (hb2_level_advance_host)
            plan.close()
            zlo = [sum(sizes[:k]) for k in range(npz)]
            boxes = [((0, 0, zlo[k]), (n, n, zlo[k] + sizes[k])) for k in range(npz)]
            lvl = abi.DeviceLevel(3, boxes, (n, n, n), flow_model=flow_model, species_gamma=gam, dx=(domain_len / n,) * 3, math=math,
                                  scheme=scheme)
            slabs = []
            for k in range(npz):
                t = torch.empty((ncomp, sizes[k] + 8, n + 8, n + 8), dtype=torch.float64).pin_memory()
                t[:, 4:-4].copy_(host[:, 4 + zlo[k]:4 + zlo[k] + sizes[k]])
                slabs.append(t)
            del host, hnp
            arrs = [t.numpy() for t in slabs]
            lvl.advance_host(arrs, dte)
            k0 = lvl.launch_count
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                lvl.advance_host(arrs, dte)
            el = time.perf_counter() - t0
            nbytes = ncomp * n * (n + 8) * (n + 8) * 8
            launches = lvl.launch_count - k0
            api = (f"hb2_level_advance_host: the host box as {npz} z-slab patches (planes per slab: {sizes}) in pinned memory; per step "
                   "H2D of every slab, 3 stages on the device-resident level as a wavefront of (slab, stage) tasks behind the uploads, "
                   "D2H of every slab as soon as it is done (full-duplex PCIe)")
            fin = all(bool(np.isfinite(a[:, 4:-4, 4:-4, 4:-4]).all()) for a in arrs[:1])
            lvl.close()
        line["e2e"] = {"value": n ** 3 * 3 * args.e2e_steps / el, "unit": "cell-updates/s",
                       "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": args.e2e_steps,
                       "size": n, "ms_per_step": 1e3 * el / args.e2e_steps, "launches": launches, "api": api}
    elif not args.no_e2e:
        # N > 1: one host box per rank (what each MPI rank of the reference owns); per step H2D of the box, ghost exchange
        # over NVLink, three stages, D2H of the new box.  Wall clock between barriers, max over ranks.
        inter = level.S[level.cur][level._interior_slices()]
        host = torch.empty(inter.shape, dtype=torch.float64).pin_memory()
        host.copy_(inter)
        torch.cuda.synchronize()
        level.rk_step_host(host, dt)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            level.rk_step_host(host, dt)
        barrier()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        nbytes = host.numel() * 8 * world
        if rank == 0:
            el = float(el[0])
            line["e2e"] = {"value": ncell_global * 3 * args.e2e_steps / el, "unit": "cell-updates/s",
                           "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": args.e2e_steps,
                           "size": list(Nglob), "ms_per_step": 1e3 * el / args.e2e_steps,
                           "api": "UniformLevel.rk_step_host: one pinned host box per rank (interior cells); per step H2D, NVLink "
                                  "ghost exchange, 3 fused stages, D2H; bytes summed over ranks"}
    elif rank == 0:
        line["e2e"] = None

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        from hamers_b200 import problems as pb
        from oracle import oracle as orc

        orc.build()
        cores = os.cpu_count() or 1
        N = args.ref_size
        model = 0 if args.model == "ss" else 1
        U, dxr, gr = (pb.convergence_single_species(3, N) if model == 0 else pb.convergence_five_eqn(3, N))
        lv = orc.PatchDesc(dim=3, n=(N,) * 3, model=model, ns=len(gr), gamma=gr, dx=dxr)
        patch = (min(32, N),) * 3
        orc.level_advance(lv, patch, U, 0.001 * dxr[0], 1, nthreads=cores)
        nst, t0 = 0, time.perf_counter()
        while True:
            orc.level_advance(lv, patch, U, 0.001 * dxr[0], 1, nthreads=cores)
            nst += 1
            el = time.perf_counter() - t0
            if el > args.cpu_seconds or nst >= 50:
                break
        line["cpu_baseline"] = {"value": N ** 3 * 3 * nst / el, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                                "sample": f"{N}^3 box in {patch[0]}^3 patches, {nst} SSP-RK3 steps, {cores} OpenMP threads, "
                                          "oracle = reference-structured CPU restatement (real reference needs SAMRAI/MPI/HDF5)"}
    elif rank == 0:
        line["cpu_baseline"] = None

    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--model", default="ss", choices=["ss", "fe"])
    ap.add_argument("--math", default="fast", choices=["fast", "exact"])
    ap.add_argument("--scheme", default="js", choices=["js", "z", "ld"],
                    help="nonlinear interpolator: WCNS5_JS (headline), WCNS5_Z, WCNS6_LD (reference-order kernels only)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--ref-size", type=int, default=128)
    ap.add_argument("--e2e-size", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-patches", type=int, default=0,
                    help="N = 1 e2e: 0 = graded z slabs (thin around the periodic seam), K > 1 = K equal slabs, 1 = one patch, unpipelined")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary measurements (Navier-Stokes, shocked field, five-eqn)")
    ap.add_argument("--ns-size", type=int, default=0, help="Navier-Stokes Taylor-Green level size (default: --size, at most 512)")
    ap.add_argument("--ns-steps", type=int, default=5)
    ap.add_argument("--shock-size", type=int, default=256)
    ap.add_argument("--fe-size", type=int, default=384)
    ap.add_argument("--no-replica", action="store_true", help="N > 1: skip the one-box replica behind parity.matches_n1")
    args = ap.parse_args()
    if args.e2e_size <= 0:
        args.e2e_size = args.size
    if args.ns_size <= 0:
        args.ns_size = min(args.size, 512)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 3)   # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
