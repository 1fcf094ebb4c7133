"""Shared helpers of the parity tests."""
from __future__ import annotations

import numpy as np

from hamers_b200 import problems as pb
from oracle import oracle as orc

# north_star tolerance: fluxes and updated conserved variables within 1e-12 relative (FP64 re-association only).
RTOL = 1.0e-12

CASES = {
    # name: (model, ns, dim, N)
    "ss2d": (orc.SINGLE_SPECIES, 1, 2, (24, 17)),
    "ss3d": (orc.SINGLE_SPECIES, 1, 3, (13, 10, 12)),
    "fe2d": (orc.FIVE_EQN_ALLAIRE, 2, 2, (24, 17)),
    "fe3d": (orc.FIVE_EQN_ALLAIRE, 2, 3, (13, 10, 12)),
}


def make_case(name, kind="random", seed=20261017):
    model, ns, dim, N = CASES[name]
    if kind == "random":
        U, dx, gam = pb.random_state(dim, N, model=model, seed=seed, shock=True)
    elif kind == "smooth":
        if model == orc.SINGLE_SPECIES:
            U, dx, gam = pb.convergence_single_species(dim, N)
        else:
            U, dx, gam = pb.convergence_five_eqn(dim, N)
    else:
        raise ValueError(kind)
    desc = orc.PatchDesc(dim=dim, n=N, model=model, ns=ns, gamma=gam, dx=dx)
    return desc, U


def rel_err(a, b):
    """max |a-b| / (|b| + field max-norm): pointwise relative error regularised by the component's scale
    (fluxes cross zero, so a purely pointwise ratio is meaningless there)."""
    a = np.asarray(a)
    b = np.asarray(b)
    axes = tuple(range(1, b.ndim))
    scale = np.abs(b).max(axis=axes, keepdims=True) if b.ndim > 1 else np.abs(b).max()
    scale = np.where(scale == 0.0, 1.0, scale)
    return float((np.abs(a - b) / (np.abs(b) + scale)).max())


def interior(desc, U, g=4):
    sl = (slice(None),) + tuple(slice(g, -g) for _ in range(desc.dim))
    return U[sl]
