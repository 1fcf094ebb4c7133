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


# Fast-arithmetic variant: FMA contraction and reciprocal sharing re-associate the reference's formulas.  On
# well-conditioned faces that is <= RTOL.  A handful of faces of the white-noise branch-coverage state are
# ILL-CONDITIONED IN THE REFERENCE'S OWN FORMULAS (e.g. the HLLC-HLL blend weights alpha_1 = |du_n|/|du|,
# alpha_2 = sqrt(1 - alpha_1^2) are built from differences of the two one-sided interpolants: when those nearly
# coincide a 1-ulp change upstream moves beta_1 by ~1e-11; the oracle itself moves by that much under 1-ulp input
# perturbations, see tests/test_oracle_conditioning.py).  Such faces are counted and bounded separately.
OUTLIER_FRACTION = 1.0e-3   # at most this share of the entries may exceed RTOL ...
OUTLIER_RTOL = 1.0e-9       # ... and none may exceed this (a wrong formula is off by >= 1e-6)


def rel_err_field(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    axes = tuple(range(1, b.ndim))
    scale = np.abs(b).max(axis=axes, keepdims=True) if b.ndim > 1 else np.abs(b).max()
    scale = np.where(scale == 0.0, 1.0, scale)
    return np.abs(a - b) / (np.abs(b) + scale)


def assert_fast_parity(a, b, what="", outlier_fraction=OUTLIER_FRACTION):
    r = rel_err_field(a, b)
    assert np.isfinite(r).all(), f"{what}: non-finite entries"
    n_out = int((r > RTOL).sum())
    assert n_out <= max(8, int(outlier_fraction * r.size)), f"{what}: {n_out} of {r.size} entries exceed {RTOL}, max {r.max():.3e}"
    assert r.max() <= OUTLIER_RTOL, f"{what}: max relative error {r.max():.3e}"
    return float(r.max()), n_out
