#!/usr/bin/env python
"""Throughput of the 2-D path (config 1 of BASELINE.json scaled up): python tools/bench2d.py [--size 8192] [--model ss|fe]
One JSON line: cell-updates/s of SSP-RK3 on a periodic N^2 level (CUDA events, 3 warm-up steps)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hamers_b200 import abi  # noqa: E402
from hamers_b200.level import UniformLevel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--model", default="ss", choices=["ss", "fe"])
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
N = args.size
model = abi.SINGLE_SPECIES if args.model == "ss" else abi.FIVE_EQN_ALLAIRE
gam = (1.4,) if args.model == "ss" else (1.6, 1.4)
lvl = UniformLevel(2, (N, N), flow_model=model, species_gamma=gam, math=abi.MATH_FAST)
xs = [torch.as_tensor(c, dtype=torch.float64, device="cuda") for c in lvl.local_coordinates()]
sin = torch.sin(np.pi * (xs[0][None, :] + xs[1][:, None]))
if args.model == "ss":
    rho = 1.0 + 0.5 * sin
    comps = [rho, rho, rho, 1.0 / 0.4 + 0.5 * rho * 2.0]
else:
    Z1 = 0.5 + 0.25 * sin
    Z2 = 1.0 - Z1
    rho = 2.0 * Z1 + Z2
    gm = 1.0 / (Z1 / 0.6 + Z2 / 0.4) + 1.0
    comps = [2.0 * Z1, Z2, rho, rho, 1.0 / (gm - 1.0) + 0.5 * rho * 2.0, Z1, Z2]
inter = lvl.interior()
for c, v in enumerate(comps):
    inter[c].copy_(v)
dt = 0.001 * lvl.dx[0]
for _ in range(3):
    lvl.rk_step(dt)
torch.cuda.synchronize()
lvl.plan.set_profiling(True)
lvl.plan.get_profile(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    lvl.rk_step(dt)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
prof = lvl.plan.get_profile(reset=True)
print(json.dumps({"workload": f"2D {args.model} periodic level {N}^2, WCNS5_JS_HLLC_HLL, SSP-RK3, fast build",
                  "value": N * N * 3 * args.steps / (ms * 1e-3), "unit": "cell-updates/s", "ms_per_step": ms / args.steps,
                  "kernels_ms": {k: round(v[0] / v[1], 3) for k, v in prof.items() if v[1] > 0},
                  "finite": bool(torch.isfinite(lvl.interior()).all())}))
