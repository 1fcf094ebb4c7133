#!/usr/bin/env python
"""Throughput of the diffusive-flux path: python tools/bench_diffusive.py [--size 256] [--dim 3] [--steps 10]
One JSON line: cells/s of hb2_compute_diffusive_flux_dev on an N^dim patch (CUDA events on the launching stream, 3 warm-up
calls, inputs resident in HBM and larger than L2 from N = 192 up in 3-D) next to its HBM roofline: ALGORITHMIC bytes per
cell = (dim + 2) conservative doubles read + dim * (dim + 2) side-flux doubles written (160 B in 3-D, 96 B in 2-D); the
primitive and node-flux scratch traffic is overhead of the three-kernel formulation and shows as achieved < peak."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hamers_b200 import abi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--dim", type=int, default=3, choices=[2, 3])
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
N, dim = args.size, args.dim
n = (N,) * dim
dx = (1.0 / N,) * dim
plan = abi.DiffusivePlan(dim, n, dx, 1.4, 2.5, 0.05, 0.02, 3.5, 0.72).use_torch_stream()
ax = (torch.arange(N + 12, dtype=torch.float64, device="cuda") - 5.5) / N
X = torch.meshgrid(*([ax] * dim), indexing="ij")
s = sum(torch.sin(2.0 * np.pi * x) for x in X)
rho = 1.0 + 0.1 * s
vel = [0.3 * torch.cos(2.0 * np.pi * X[a]) for a in range(dim)]
E = 2.5 + 0.5 * rho * sum(v * v for v in vel)
Q = torch.stack([rho] + [rho * v for v in vel] + [E]).contiguous()
F = [torch.empty((dim + 2,) + plan.side_shape(a), dtype=torch.float64, device="cuda") for a in range(dim)]
dt = 1.0e-4
for _ in range(3):
    plan.compute_diffusive_flux(Q, dt, F)
torch.cuda.synchronize()
l0 = plan.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    plan.compute_diffusive_flux(Q, dt, F)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
cells = float(N) ** dim
alg_bytes = cells * 8.0 * ((dim + 2) + dim * (dim + 2))
peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
        peak, peak_src = float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
except Exception:
    pass
ach = alg_bytes / (ms * 1e-3) / 1e9
print(json.dumps({"workload": f"{dim}D single-species diffusive flux (SIXTH_ORDER), {N}^{dim} patch, exact arithmetic",
                  "value": cells / (ms * 1e-3), "unit": "cells/s", "ms_per_call": ms,
                  "gpu_launches": plan.launch_count - l0,
                  "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                               "traffic": None},
                  "finite": bool(all(torch.isfinite(f).all() for f in F))}))
plan.close()
