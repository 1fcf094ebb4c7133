"""Initial conditions and error norms of the reference's convergence problems (host-side, numpy).

These are the fixtures of BASELINE.json configs 1-3:
  * single-species density-wave advection, 2D/3D:
      problems/Euler/initial_conditions/ConvergenceSingleSpecies.cpp:77-99 (2D), :145-174 (3D)
      exact solution / norms: problems/Euler/error_statistics/ConvergenceSingleSpecies.cpp:85, :199-254
  * five-equation Allaire volume-fraction wave, 2D/3D:
      problems/Euler/initial_conditions/ConvergenceFiveEqnAllaire.cpp:176-219
      error on Z_1: problems/Euler/error_statistics/ConvergenceFiveEqnAllaire.cpp:198-201
Domain [-1,1]^d, fully periodic (tests/*/convergence_test.py:102).

Arrays are (ncomp, [z,] y, x) with x fastest, i.e. the SAMRAI column-major CellData layout
seen from numpy; they hold the level INTERIOR only.
"""
from __future__ import annotations

import numpy as np

SINGLE_SPECIES = 0
FIVE_EQN_ALLAIRE = 1


def cell_centres(dim: int, N, xlo=-1.0, xhi=1.0):
    """Cell-centre coordinates x = xlo + (i + 1/2)*dx per axis; returns (list of 1-D arrays, dx tuple)."""
    N = tuple(N) if np.iterable(N) else (int(N),) * dim
    dx = tuple((xhi - xlo) / n for n in N)
    xs = [xlo + (np.arange(n, dtype=np.float64) + 0.5) * h for n, h in zip(N, dx)]
    return xs, dx


def _sum_coords(dim, xs):
    # s = x + y (+ z), summed in the reference's order, broadcast to ([z,] y, x)
    if dim == 2:
        return xs[0][None, :] + xs[1][:, None]
    return (xs[0][None, None, :] + xs[1][None, :, None]) + xs[2][:, None, None]


def convergence_single_species(dim: int, N):
    """Returns (U, dx, gamma) with U = [rho, rho*u, rho*v, (rho*w), E]; u=v=w=1, p=1, gamma=7/5."""
    xs, dx = cell_centres(dim, N)
    s = _sum_coords(dim, xs)
    gamma = 7.0 / 5.0
    rho = 1.0 + 0.5 * np.sin(np.pi * s)
    u = 1.0
    p = 1.0
    ke = float(dim)  # u*u + v*v (+ w*w)
    E = p / (gamma - 1.0) + 0.5 * rho * ke
    U = np.stack([rho] + [rho * u] * dim + [E]).astype(np.float64)
    return np.ascontiguousarray(U), dx, (gamma,)


def exact_density_single_species(dim: int, N, time: float):
    xs, _ = cell_centres(dim, N)
    s = _sum_coords(dim, xs)
    return 1.0 + 0.5 * np.sin(np.pi * (s - float(dim) * time))


def convergence_five_eqn(dim: int, N):
    """Returns (U, dx, gammas) with U = [Zrho_1, Zrho_2, rho*u, rho*v, (rho*w), E, Z_1, Z_2]."""
    xs, dx = cell_centres(dim, N)
    s = _sum_coords(dim, xs)
    g1, g2 = 8.0 / 5.0, 7.0 / 5.0
    r1, r2 = 2.0, 1.0
    Z1 = 0.5 + 0.25 * np.sin(np.pi * s)
    Z2 = 1.0 - Z1
    Zr1 = Z1 * r1
    Zr2 = Z2 * r2
    rho_m = Zr1 + Zr2
    gamma_m = 1.0 / (Z1 / (g1 - 1.0) + Z2 / (g2 - 1.0)) + 1.0
    p = 1.0
    E = p / (gamma_m - 1.0) + 0.5 * rho_m * float(dim)
    U = np.stack([Zr1, Zr2] + [rho_m * 1.0] * dim + [E, Z1, Z2]).astype(np.float64)
    return np.ascontiguousarray(U), dx, (g1, g2)


def exact_Z1_five_eqn(dim: int, N, time: float):
    xs, _ = cell_centres(dim, N)
    s = _sum_coords(dim, xs)
    return 0.5 + 0.25 * np.sin(np.pi * (s - float(dim) * time))


def error_norms(numerical: np.ndarray, exact: np.ndarray, dx):
    """Volume-weighted L1, L2 and max norms over level 0
    (problems/Euler/error_statistics/ConvergenceSingleSpecies.cpp:203-249)."""
    dvol = float(np.prod(dx))
    err = np.abs(exact - numerical)
    vol = dvol * err.size
    L1 = float((dvol * err).sum() / vol)
    L2 = float(np.sqrt((dvol * err * err).sum() / vol))
    Linf = float(err.max())
    return L1, L2, Linf


def pad_periodic(U: np.ndarray, g: int = 4):
    """Level interior -> one ghost box with periodic images (what a same-level periodic
    xfer::RefineSchedule::fillData produces for a single patch covering the level)."""
    pad = [(0, 0)] + [(g, g)] * (U.ndim - 1)
    return np.ascontiguousarray(np.pad(U, pad, mode="wrap"))


def random_state(dim: int, N, model=SINGLE_SPECIES, seed=20261017, shock=True):
    """Branch-coverage input M2: randomised positive state plus a planar
    Mach-3 shock slab that trips the sensor, the HLL upwind overrides and the first-order
    fallback.  Returns (U, dx, gammas) on the level interior."""
    rng = np.random.default_rng(seed)
    N = tuple(N) if np.iterable(N) else (int(N),) * dim
    shape = tuple(reversed(N))
    dx = tuple(2.0 / n for n in N)
    rho = rng.uniform(0.5, 2.0, shape)
    vel = [rng.uniform(-1.0, 1.0, shape) for _ in range(dim)]
    p = rng.uniform(0.5, 2.0, shape)
    if shock:
        # slab normal to x in the middle third: post-shock state of a Mach-3 shock (gamma = 1.4)
        n0 = N[0]
        sl = (Ellipsis, slice(n0 // 3, 2 * n0 // 3))
        rho[sl] = 3.857143 * (1.0 + 0.01 * rng.standard_normal(rho[sl].shape))
        p[sl] = 10.33333 * (1.0 + 0.01 * rng.standard_normal(rho[sl].shape))
        vel[0][sl] = 2.629369
        # a supersonic stream to hit s_L > 0 / s_R < 0
        sl2 = (Ellipsis, slice(0, max(2, n0 // 8)))
        vel[0][sl2] = 4.0
        sl3 = (Ellipsis, slice(n0 - max(2, n0 // 8), n0))
        vel[0][sl3] = -4.0
    ke = sum(v * v for v in vel)
    if model == SINGLE_SPECIES:
        gam = (1.4,)
        E = p / (gam[0] - 1.0) + 0.5 * rho * ke
        U = np.stack([rho] + [rho * v for v in vel] + [E])
    else:
        gam = (1.6, 1.4)
        Z1 = rng.uniform(0.05, 0.95, shape)
        Z2 = 1.0 - Z1
        r1 = rho * rng.uniform(0.8, 1.2, shape)
        r2 = rho * rng.uniform(0.4, 0.8, shape)
        Zr1, Zr2 = Z1 * r1, Z2 * r2
        rho_m = Zr1 + Zr2
        gamma_m = 1.0 / (Z1 / (gam[0] - 1.0) + Z2 / (gam[1] - 1.0)) + 1.0
        E = p / (gamma_m - 1.0) + 0.5 * rho_m * ke
        U = np.stack([Zr1, Zr2] + [rho_m * v for v in vel] + [E, Z1, Z2])
    return np.ascontiguousarray(U.astype(np.float64)), dx, gam
