#!/usr/bin/env python
"""Small single-process run of the hot path for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_case.py [ss|fe] [fast|exact]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hamers_b200 import abi  # noqa: E402
from hamers_b200 import problems as pb  # noqa: E402
from hamers_b200.level import UniformLevel  # noqa: E402

model = abi.FIVE_EQN_ALLAIRE if (len(sys.argv) > 1 and sys.argv[1] == "fe") else abi.SINGLE_SPECIES
math = abi.MATH_EXACT if (len(sys.argv) > 2 and sys.argv[2] == "exact") else abi.MATH_FAST
N = (40, 37, 44)
U, dx, gam = pb.random_state(3, N, model=model, seed=2, shock=True)
for push in (False, True):
    lvl = UniformLevel(3, N, flow_model=model, species_gamma=gam, math=math, push=push)
    lvl.set_interior(U)
    lvl.advance(2.0e-4, 1)
    torch.cuda.synchronize()
    print("push", push, "finite", bool(torch.isfinite(lvl.interior()).all()), "dt", lvl.stable_dt(0.5))
    lvl.close()
# the API-preserving mode (materialised side fluxes)
plan = abi.Plan(3, N, flow_model=model, species_gamma=gam, dx=dx, math=math).use_torch_stream()
Q = torch.from_numpy(pb.pad_periodic(U)).cuda()
neq = plan.neq
F = [torch.empty((neq,) + plan.side_shape(a), dtype=torch.float64, device="cuda") for a in range(3)]
S = torch.zeros((neq,) + plan.cell_shape, dtype=torch.float64, device="cuda")
plan.compute_flux_and_source(Q, 1e-3, F, S)
torch.cuda.synchronize()
print("emit finite", all(bool(torch.isfinite(f).all()) for f in F))
plan.close()
