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


# ---- FOUR_EQN_CONSERVATIVE flow model (partial densities rho Y_i, momentum, total energy) ------------------
FOUR_EQN_CONSERVATIVE = 2


def mixture_gamma_mass_fractions(Y, gam, R):
    """gamma = sum Y_i c_p_i / sum Y_i c_v_i (EquationOfStateMixingRulesIdealGas.cpp:4677-4738)."""
    c_p = sum(y * g / (g - 1.0) * r for y, g, r in zip(Y, gam, R))
    c_v = sum(y * 1.0 / (g - 1.0) * r for y, g, r in zip(Y, gam, R))
    return c_p / c_v


def convergence_four_eqn(dim: int, N):
    """Two-species analogue of the reference's convergence problems for the four-eqn conservative model (the reference
    ships none for this model): partial-density waves advected with u = v = (w =) 1 at uniform pressure 1.
    Returns (U, dx, gammas, Rs) with U = [rhoY_1, rhoY_2, rho*u, rho*v, (rho*w), E]."""
    xs, dx = cell_centres(dim, N)
    s = _sum_coords(dim, xs)
    gam, R = (8.0 / 5.0, 7.0 / 5.0), (1.0, 2.0)
    rY1 = 1.0 + 0.5 * np.sin(np.pi * s)
    rY2 = 1.0 - 0.25 * np.sin(np.pi * s)
    rho = rY1 + rY2
    g = mixture_gamma_mass_fractions((rY1 / rho, rY2 / rho), gam, R)
    E = 1.0 / (g - 1.0) + 0.5 * rho * float(dim)
    U = np.stack([rY1, rY2] + [rho * 1.0] * dim + [E]).astype(np.float64)
    return np.ascontiguousarray(U), dx, gam, R


def exact_rhoY1_four_eqn(dim: int, N, time: float):
    xs, _ = cell_centres(dim, N)
    s = _sum_coords(dim, xs)
    return 1.0 + 0.5 * np.sin(np.pi * (s - float(dim) * time))


def random_state_four_eqn(dim: int, N, seed=20261017, shock=True):
    """The branch-coverage state of random_state with two species of different gamma and R.  Returns (U, dx, gammas, Rs)."""
    U1, dx, _ = random_state(dim, N, model=SINGLE_SPECIES, seed=seed, shock=shock)
    rng = np.random.default_rng(seed + 1)
    gam, R = (1.6, 1.4), (0.7, 1.3)
    rho = U1[0]
    Y1 = rng.uniform(0.02, 0.98, rho.shape)
    vel = [U1[1 + a] / rho for a in range(dim)]
    p = 0.4 * (U1[dim + 1] - 0.5 * rho * sum(v * v for v in vel))
    g = mixture_gamma_mass_fractions((Y1, 1.0 - Y1), gam, R)
    E = p / (g - 1.0) + 0.5 * rho * sum(v * v for v in vel)
    U = np.stack([rho * Y1, rho * (1.0 - Y1)] + [rho * v for v in vel] + [E])
    return np.ascontiguousarray(U.astype(np.float64)), dx, gam, R


def random_state_three_species(dim: int, N, model, seed=20261017, shock=True):
    """The branch-coverage state of random_state with THREE species (the reference's flow models are generic in
    d_num_species; its shipped decks use two).  model = FIVE_EQN_ALLAIRE: U = [Z1 rho1, Z2 rho2, Z3 rho3, rho u.., E, Z1, Z2,
    Z3] (the last volume fraction is the stored, derived component); returns (U, dx, gammas).  Any other model value is taken
    as FOUR_EQN_CONSERVATIVE: U = [rho Y1, rho Y2, rho Y3, rho u.., E]; returns (U, dx, gammas, Rs)."""
    U1, dx, _ = random_state(dim, N, model=SINGLE_SPECIES, seed=seed, shock=shock)
    rng = np.random.default_rng(seed + 2)
    gam = (1.6, 1.4, 1.25)
    rho = U1[0]
    vel = [U1[1 + a] / rho for a in range(dim)]
    ke = sum(v * v for v in vel)
    p = 0.4 * (U1[dim + 1] - 0.5 * rho * ke)
    w = rng.uniform(0.05, 1.0, (3,) + rho.shape)
    frac = w / w.sum(axis=0)                       # three fractions in (0, 1) that sum to one
    if model == FIVE_EQN_ALLAIRE:
        Z = [frac[0], frac[1], 1.0 - frac[0] - frac[1]]
        r = [rho * rng.uniform(0.8, 1.2, rho.shape), rho * rng.uniform(0.4, 0.8, rho.shape), rho * rng.uniform(1.1, 1.5, rho.shape)]
        Zr = [Z[i] * r[i] for i in range(3)]
        rho_m = Zr[0] + Zr[1] + Zr[2]
        gamma_m = 1.0 / sum(Z[i] / (gam[i] - 1.0) for i in range(3)) + 1.0
        E = p / (gamma_m - 1.0) + 0.5 * rho_m * ke
        U = np.stack(Zr + [rho_m * v for v in vel] + [E] + Z)
        return np.ascontiguousarray(U.astype(np.float64)), dx, gam
    R = (0.7, 1.3, 1.0)
    Y = [frac[0], frac[1], 1.0 - frac[0] - frac[1]]
    g = mixture_gamma_mass_fractions(Y, gam, R)
    E = p / (g - 1.0) + 0.5 * rho * ke
    U = np.stack([rho * y for y in Y] + [rho * v for v in vel] + [E])
    return np.ascontiguousarray(U.astype(np.float64)), dx, gam, R


# species 0: SF6, species 1: air (input_2D_Richtmyer_Meshkov_instability.txt:21-22)
RMI_GAMMA = (1.09312, 1.39909)
RMI_R = (56.927, 296.803)


def richtmyer_meshkov_2d(N, x_up=(0.016, 0.001)):
    """problems/Euler/initial_conditions/RichtmyerMeshkovInstability2D.cpp:58-160 on N = (Nx, Ny) cells of the deck's domain
    [0, 0.016] x [0, 0.001] (BASELINE.json config 4).  Returns (U, dx, gammas, Rs), U = [rhoY_SF6, rhoY_air, rho u, rho v, E]."""
    from math import erf

    Nx, Ny = N
    dx = (x_up[0] / Nx, x_up[1] / Ny)
    x = (np.arange(Nx) + 0.5) * dx[0]
    y = (np.arange(Ny) + 0.5) * dx[1]
    X, Yc = np.meshgrid(x, y, indexing="xy")          # shape (Ny, Nx), x fastest
    D = 0.001
    eps_i = 6.0 / 128.0 * D
    g1 = 1.39909
    c_p = (668.286, 1040.50)
    c_v = (611.359, 743.697)
    rho_SF6, u_SF6, p_SF6 = 5.972856, 436.201332, 101325.0
    rho_pre, u_pre, p_pre = 1.145598, 436.201332, 101325.0
    rho_post, u_post, p_post = 1.616874, 309.060123, 164859.0
    dR = X - (2.0 / 5.0 - 1.0 / 10.0 * np.sin(2 * np.pi * (Yc / D + 1.0 / 4.0))) * D
    f_sm = 0.5 * (1.0 + np.vectorize(erf)(dR / eps_i))
    rY0 = rho_SF6 * (1.0 - f_sm)
    rY1 = rho_pre * f_sm
    u_i = u_SF6 * (1.0 - f_sm) + u_pre * f_sm
    p_i = p_SF6 * (1.0 - f_sm) + p_pre * f_sm
    rho_i = rY0 + rY1
    Y0 = rY0 / rho_i
    Y1 = 1.0 - Y0
    gamma = (Y0 * c_p[0] + Y1 * c_p[1]) / (Y0 * c_v[0] + Y1 * c_v[1])
    E = p_i / (gamma - 1.0) + 0.5 * rho_i * (u_i * u_i)
    ru = rho_i * u_i
    post = X > 7.0 / 10.0 * D
    rY0 = np.where(post, 0.0, rY0)
    rY1 = np.where(post, rho_post, rY1)
    ru = np.where(post, rho_post * u_post, ru)
    E = np.where(post, p_post / (g1 - 1.0) + 0.5 * rho_post * u_post * u_post, E)
    U = np.stack([rY0, rY1, ru, np.zeros_like(ru), E]).astype(np.float64)
    return np.ascontiguousarray(U), dx, RMI_GAMMA, RMI_R
