#!/usr/bin/env python
"""End-to-end step on host memory through hb2_level_advance_host with a chosen stack of z slabs:
    python tools/bench_e2e.py [--size 512] [--sizes 16,72,72,72,72,72,72,16,16,16,16] [--steps 3]
    HB2_LEVEL_TRACE=1 python tools/bench_e2e.py ...     # per-patch timeline of the last step on stderr
One line per run: slab sizes, ms per step, cell-updates/s (wall clock around the calls; pinned host slabs)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hamers_b200 import abi  # noqa: E402
from hamers_b200 import problems as pb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--sizes", action="append", default=[], help="comma-separated planes per slab (repeatable)")
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
n = args.size
U, _, gam = pb.convergence_single_species(3, 64)
tile = n // 64
U = np.tile(U, (1, tile, tile, tile))
for spec in args.sizes or [",".join(str(x) for x in [n // 32] + [(n - 5 * (n // 32)) // 6] * 6 + [n // 32] * 4)]:
    sizes = [int(x) for x in spec.split(",")]
    assert sum(sizes) == n, (sum(sizes), n)
    zlo = [sum(sizes[:k]) for k in range(len(sizes))]
    boxes = [((0, 0, zlo[k]), (n, n, zlo[k] + sizes[k])) for k in range(len(sizes))]
    lvl = abi.DeviceLevel(3, boxes, (n, n, n), flow_model=abi.SINGLE_SPECIES, species_gamma=(1.4,), dx=(2.0 / n,) * 3, math=abi.MATH_FAST)
    slabs = []
    for k in range(len(sizes)):
        t = torch.zeros((5, sizes[k] + 8, n + 8, n + 8), dtype=torch.float64).pin_memory()
        t[:, 4:-4, 4:-4, 4:-4].copy_(torch.from_numpy(U[:, zlo[k]:zlo[k] + sizes[k]]))
        slabs.append(t)
    arrs = [t.numpy() for t in slabs]
    dt = 0.001 * 2.0 / n
    lvl.advance_host(arrs, dt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lvl.advance_host(arrs, dt)
    el = (time.perf_counter() - t0) / args.steps
    # the same stage loop on the device-resident level (no transfers): what the slab decomposition costs the kernels
    lvl.upload(arrs)
    lvl.advance(dt)
    lvl.synchronize()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        lvl.advance(dt)
    lvl.synchronize()
    dev = (time.perf_counter() - t1) / args.steps
    print(f"sizes {spec}  device_resident_ms_per_step {1e3 * dev:.2f}  ms_per_step {1e3 * el:.2f}  cell_updates_per_s {n ** 3 * 3 / el:.4e}  finite {bool(np.isfinite(arrs[0]).all())}", flush=True)
    lvl.close()
    del slabs, arrs
