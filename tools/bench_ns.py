#!/usr/bin/env python
"""Throughput of the single-species Navier-Stokes level (config 5 of BASELINE.json made viscous):
python tools/bench_ns.py [--size 256] [--steps 5] [--math 1]
3-D Taylor-Green vortex on a periodic N^3 level, WCNS5_JS_HLLC_HLL + SIXTH_ORDER diffusive flux, SSP-RK3.  One JSON line:
cell-updates/s (N^3 x 3 stages per step; CUDA events on the launching stream, 3 warm-up steps, state larger than L2 from
N = 192 up), the launches per step and the share of the step spent in the diffusive calls (second timing pass)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hamers_b200 import abi  # noqa: E402
from hamers_b200.ns_level import NavierStokesLevel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--math", type=int, default=abi.MATH_FAST)
args = ap.parse_args()
N = args.size
two_pi = 2.0 * np.pi
lvl = NavierStokesLevel(3, (N, N, N), species_gamma=1.4, species_R=1.0, species_mu=1.0 / 1600.0, species_mu_v=0.0,
                        species_c_p=3.5, species_Pr=0.71, domain=(0.0, two_pi), math=args.math)
x, y, z = [torch.as_tensor(c, dtype=torch.float64, device="cuda") for c in lvl.coordinates()]
X, Y, Z = x[None, None, :], y[None, :, None], z[:, None, None]
M0 = 0.1                                       # Mach number of the vortex
u = torch.sin(X) * torch.cos(Y) * torch.cos(Z)
v = -torch.cos(X) * torch.sin(Y) * torch.cos(Z)
w = torch.zeros_like(u)
p = 1.0 / (1.4 * M0 * M0) + (torch.cos(2 * X) + torch.cos(2 * Y)) * (torch.cos(2 * Z) + 2.0) / 16.0
rho = torch.ones_like(u)
E = p / 0.4 + 0.5 * rho * (u * u + v * v + w * w)
inter = lvl.interior()
for c, f in enumerate([rho, rho * u, rho * v, rho * w, E]):
    inter[c].copy_(f.expand_as(inter[c]))
dt = 0.2 * lvl.dx[0] / (1.0 + 1.0 / M0)
for _ in range(3):
    lvl.rk_step(dt)
torch.cuda.synchronize()
l0 = lvl.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    lvl.rk_step(dt)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
launches = (lvl.launch_count - l0) / args.steps
# the diffusive part of the stages alone, same state (fast route: flux-free divergence update; exact: side fluxes)
scratch = torch.zeros_like(lvl.S[lvl.cur])
Fd = [torch.empty((5,) + lvl.dplan.side_shape(a), dtype=torch.float64, device="cuda") for a in range(3)]
d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d0.record()
for _ in range(3 * args.steps):
    if args.math == abi.MATH_EXACT:
        lvl.dplan.compute_diffusive_flux(lvl.S[lvl.cur], dt, Fd)
    else:
        lvl.dplan.divergence_accumulate(lvl.S[lvl.cur], dt, 6, 1.0, scratch)
d1.record()
torch.cuda.synchronize()
ms_diff = d0.elapsed_time(d1) / args.steps
print(json.dumps({"workload": f"3D single-species Navier-Stokes, Taylor-Green vortex Re=1600 M=0.1, periodic {N}^3, "
                              f"WCNS5_JS_HLLC_HLL + SIXTH_ORDER, SSP-RK3, {'fast' if args.math else 'exact'} build",
                  "value": float(N) ** 3 * 3 / (ms * 1e-3), "unit": "cell-updates/s", "ms_per_step": ms,
                  "gpu_launches_per_step": launches, "diffusive_ms_per_step": ms_diff,
                  "diffusive_share": ms_diff / ms, "finite": bool(torch.isfinite(lvl.interior()).all())}))
lvl.close()
