"""ctypes front end of the CPU ORACLE (test infrastructure, NOT a product path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (hamers_b200) never does.

The C sources restate the reference's algorithm (see hamers_oracle.h for citations and the
parity status).  This module only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

G = 4
MAX_SPECIES = 4
SINGLE_SPECIES = 0
FIVE_EQN_ALLAIRE = 1
FOUR_EQN_CONSERVATIVE = 2

# SSP-RK3 default table, RungeKuttaLevelIntegrator.cpp:3894-3929
SSPRK3_ALPHA = np.array([[1.0, 0.0, 0.0], [3.0 / 4.0, 1.0 / 4.0, 0.0], [1.0 / 3.0, 0.0, 2.0 / 3.0]])
SSPRK3_BETA = np.array([[1.0, 0.0, 0.0], [0.0, 1.0 / 4.0, 0.0], [0.0, 0.0, 2.0 / 3.0]])
# weights of the flux / source sums of a whole step (RungeKuttaLevelIntegrator.cpp:3894-3929), used by the AMR flux correction
SSPRK3_GAMMA = np.array([[1.0 / 6.0, 0.0, 0.0], [0.0, 1.0 / 6.0, 0.0], [0.0, 0.0, 2.0 / 3.0]])


class _Desc(C.Structure):
    _fields_ = [
        ("dim", C.c_int),
        ("n", C.c_int * 3),
        ("model", C.c_int),
        ("ns", C.c_int),
        ("gamma", C.c_double * MAX_SPECIES),
        ("dx", C.c_double * 3),
        ("weno_p", C.c_int),
        ("scheme", C.c_int),
        ("weno_q", C.c_int),
        ("weno_C", C.c_double),
        ("weno_alpha_tau", C.c_double),
        ("R", C.c_double * MAX_SPECIES),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the reference's CI flags (gcc -O3, no SIMD pragmas, no FMA)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("hamers_oracle.c", "oracle_level.c", "hamers_oracle.h", "oracle_diffusive.c",
                                                   "oracle_diffusive.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_num_eqn.argtypes = [C.POINTER(_Desc)]
        _LIB.orc_num_comp.argtypes = [C.POINTER(_Desc)]
        _LIB.orc_compute_flux_and_source.restype = C.c_int
        _LIB.orc_advance_stage.restype = C.c_int
        _LIB.orc_level_advance.restype = C.c_int
    return _LIB


@dataclass
class PatchDesc:
    """Mirror of orc_desc: one patch (or one level) of the hot path."""
    dim: int
    n: tuple
    model: int = SINGLE_SPECIES
    ns: int = 1
    gamma: tuple = (1.4,)
    dx: tuple = (1.0, 1.0, 1.0)
    weno_p: int = 2
    scheme: int = 0            # 0 WCNS5-JS, 1 WCNS5-Z, 2 WCNS6-LD
    weno_q: int = 4
    weno_C: float = 1.0e9
    weno_alpha_tau: float = 35.0
    R: tuple = ()              # four-eqn conservative: species gas constants (species_R)
    _c: _Desc = field(default=None, repr=False)

    def c(self) -> _Desc:
        d = _Desc()
        d.dim = self.dim
        for a in range(3):
            d.n[a] = int(self.n[a]) if a < self.dim else 1
            d.dx[a] = float(self.dx[a]) if a < self.dim else 1.0
        d.model = self.model
        d.ns = self.ns
        for i, g in enumerate(self.gamma):
            d.gamma[i] = float(g)
        d.weno_p = self.weno_p
        d.scheme, d.weno_q, d.weno_C, d.weno_alpha_tau = self.scheme, self.weno_q, self.weno_C, self.weno_alpha_tau
        for i, r in enumerate(self.R):
            d.R[i] = float(r)
        return d

    @property
    def neq(self) -> int:
        if self.model == FOUR_EQN_CONSERVATIVE:
            return self.dim + 1 + self.ns
        return self.dim + 2 if self.model == SINGLE_SPECIES else self.dim + 2 * self.ns

    @property
    def ncomp(self) -> int:
        return self.neq + 1 if self.model == FIVE_EQN_ALLAIRE else self.neq

    @property
    def ghost_shape(self):
        """numpy shape (z, y, x) of one ghost-box component (x fastest, SAMRAI column-major)."""
        return tuple(int(self.n[a]) + 2 * G for a in reversed(range(self.dim)))

    @property
    def cell_shape(self):
        return tuple(int(self.n[a]) for a in reversed(range(self.dim)))

    def side_shape(self, direction: int):
        e = [int(self.n[a]) for a in range(self.dim)]
        e[direction] += 1
        return tuple(reversed(e))

    def mid_shape(self, direction: int):
        e = [int(self.n[a]) for a in range(self.dim)]
        e[direction] += 3
        return tuple(reversed(e))


def _pp(arrs):
    """array of double* from a list of contiguous float64 arrays (None -> NULL)."""
    P = (C.POINTER(C.c_double) * len(arrs))()
    for i, a in enumerate(arrs):
        if a is not None:
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
            P[i] = a.ctypes.data_as(C.POINTER(C.c_double))
    return P


def compute_flux_and_source(desc: PatchDesc, Q: np.ndarray, dt: float, source=None, debug=False):
    """Q: (ncomp, *ghost_shape).  Returns (F list per direction of (neq, *side_shape), S (neq, *cell_shape))
    and, with debug=True, additionally (F_mid list, sensor list)."""
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    assert Q.shape == (desc.ncomp,) + desc.ghost_shape, (Q.shape, desc.ghost_shape)
    neq, dim = desc.neq, desc.dim
    F = [np.full((neq,) + desc.side_shape(a), np.nan) for a in range(dim)]
    S = np.zeros((neq,) + desc.cell_shape) if source is None else source
    Fm = [np.full((neq,) + desc.mid_shape(a), np.nan) for a in range(dim)] if debug else None
    sen = [np.full(desc.mid_shape(a), np.nan) for a in range(dim)] if debug else None
    d = desc.c()
    Fp = _pp([F[a][e] for a in range(dim) for e in range(neq)])
    Sp = _pp([S[e] for e in range(neq)])
    Qp = _pp([Q[c] for c in range(desc.ncomp)])
    Fmp = _pp([Fm[a][e] for a in range(dim) for e in range(neq)]) if debug else None
    sp = _pp(sen) if debug else None
    rc = lib().orc_compute_flux_and_source(C.byref(d), Qp, C.c_double(dt), Fp, Sp, Fmp, sp)
    assert rc == 0
    if debug:
        return F, S, Fm, sen
    return F, S


def advance_stage(desc: PatchDesc, alpha, beta, U_int, F_int, S_int):
    """One Euler::advanceSingleStepOnPatch.  U_int: list of (ncomp,*ghost_shape); F_int: list (per m) of
    lists (per dir) of (neq,*side_shape) or None; S_int likewise.  Returns U_out (ncomp,*ghost_shape)."""
    ncoef = len(alpha)
    neq, dim = desc.neq, desc.dim
    U_out = np.zeros((desc.ncomp,) + desc.ghost_shape)
    d = desc.c()
    keep = []
    PP = C.POINTER(C.POINTER(C.c_double))
    Ut = (PP * ncoef)()
    Ft = (PP * ncoef)()
    St = (PP * ncoef)()
    for m in range(ncoef):
        Um = np.ascontiguousarray(U_int[m])
        up = _pp([Um[c] for c in range(desc.ncomp)])
        if F_int[m] is not None:
            fp = _pp([F_int[m][a][e] for a in range(dim) for e in range(neq)])
            sp = _pp([S_int[m][e] for e in range(neq)])
        else:
            fp = _pp([None] * (dim * neq))
            sp = _pp([None] * neq)
        keep += [Um, up, fp, sp]
        Ut[m] = C.cast(up, PP)
        Ft[m] = C.cast(fp, PP)
        St[m] = C.cast(sp, PP)
    a = (C.c_double * ncoef)(*[float(x) for x in alpha])
    b = (C.c_double * ncoef)(*[float(x) for x in beta])
    rc = lib().orc_advance_stage(C.byref(d), ncoef, a, b, Ut, Ft, St, _pp([U_out[c] for c in range(desc.ncomp)]))
    assert rc == 0
    return U_out


def level_advance(level: PatchDesc, patch, U: np.ndarray, dt: float, nsteps: int,
                  alpha=SSPRK3_ALPHA, beta=SSPRK3_BETA, nthreads: int = 0):
    """Advance a periodic uniform level in place.  U: (ncomp, *cell_shape of the level)."""
    assert U.dtype == np.float64 and U.flags["C_CONTIGUOUS"]
    assert U.shape == (level.ncomp,) + level.cell_shape
    nst = alpha.shape[0]
    d = level.c()
    p = (C.c_int * 3)(*[int(patch[a]) if a < level.dim else 1 for a in range(3)])
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    b = np.ascontiguousarray(beta, dtype=np.float64)
    rc = lib().orc_level_advance(C.byref(d), p, _pp([U[c] for c in range(level.ncomp)]), C.c_double(dt),
                                 int(nsteps), int(nst), a.ctypes.data_as(C.POINTER(C.c_double)),
                                 b.ctypes.data_as(C.POINTER(C.c_double)), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"orc_level_advance failed: {rc}")
    return U


def spectral_radii_and_dt(desc: PatchDesc, Q: np.ndarray, include_ghosts: bool = True):
    """Euler::computeSpectralRadiusesAndStableDtOnPatch: (spectral radii per direction, stable dt for CFL = 1)."""
    out = (C.c_double * 4)()
    d = desc.c()
    rc = lib().orc_spectral_radii_and_dt(C.byref(d), _pp([Q[c] for c in range(desc.ncomp)]), 1 if include_ghosts else 0, out)
    assert rc == 0
    return np.array(out[:desc.dim]), float(out[desc.dim])


def path_points(vals):
    """orc_path_points: (derivative, theta, Omega, sensor value, face flux) from 16 inputs."""
    a = (C.c_double * 16)(*[float(x) for x in vals])
    out = (C.c_double * 5)()
    lib().orc_path_points(a, out)
    return list(out)


def path_points2(vals):
    """orc_path_points2: derived data, characteristic projection / back-projection, RK update from 32 inputs."""
    a = (C.c_double * 32)(*[float(x) for x in vals])
    out = (C.c_double * 20)()
    lib().orc_path_points2(a, out)
    return list(out)


def path_points3(vals):
    """orc_path_points3: five-eqn derived data, projection / back-projection, advective source from 56 inputs."""
    a = (C.c_double * 56)(*[float(x) for x in vals])
    out = (C.c_double * 32)()
    lib().orc_path_points3(a, out)
    return list(out)


def path_points4(vals):
    """orc_path_points4: five-eqn (rho, c, epsilon, 0) of one interpolated side from V[7] and two species gammas."""
    a = (C.c_double * 12)(*([float(x) for x in vals] + [0.0] * (12 - len(vals))))
    out = (C.c_double * 4)()
    lib().orc_path_points4(a, out)
    return list(out)


def path_points5(vals):
    """orc_path_points5: bounds flags (five-eqn side V[7], gammas, direction | single-species side V[5])."""
    a = (C.c_double * 16)(*([float(x) for x in vals] + [0.0] * (16 - len(vals))))
    out = (C.c_double * 2)()
    lib().orc_path_points5(a, out)
    return list(out)


def path_points7(vals):
    """orc_path_points7: four-eqn conservative mixture chain of a cell, side state, bounds flag."""
    a = (C.c_double * 16)(*[float(x) for x in vals])
    out = (C.c_double * 15)()
    lib().orc_path_points7(a, out)
    return list(out)


def path_points8(vals):
    """orc_path_points8: four-eqn conservative face averages, projection, back-projection (3-D x)."""
    a = (C.c_double * 24)(*([float(x) for x in vals] + [0.0] * (24 - len(vals))))
    out = (C.c_double * 16)()
    lib().orc_path_points8(a, out)
    return list(out)


def _neq_of(model, dim, ns):
    if model == FOUR_EQN_CONSERVATIVE:
        return dim + 1 + ns
    return dim + 2 if model == SINGLE_SPECIES else dim + 2 * ns


def path_points6(vals):
    """orc_path_points6: spectral radii of one cell, their sum, running maximum, stable dt."""
    a = (C.c_double * 8)(*[float(x) for x in vals])
    out = (C.c_double * 6)()
    lib().orc_path_points6(a, out)
    return list(out)


def constants():
    out = (C.c_double * 7)()
    lib().orc_constants(out)
    return list(out)


def eos_point(gamma, rho, epsilon):
    """(p, c, epsilon recovered from p) of the ideal-gas EOS as the oracle's path evaluates them."""
    p, c, e = C.c_double(), C.c_double(), C.c_double()
    lib().orc_eos_point(C.c_double(gamma), C.c_double(rho), C.c_double(epsilon), C.byref(p), C.byref(c), C.byref(e))
    return p.value, c.value, e.value


def weno5js_point(U, p=2):
    Ua = (C.c_double * 6)(*[float(x) for x in U])
    m, pl = C.c_double(), C.c_double()
    lib().orc_weno5js_point(Ua, int(p), C.byref(m), C.byref(pl))
    return m.value, pl.value


def weno5z_point(U, p=2):
    Ua = (C.c_double * 6)(*[float(x) for x in U])
    m, pl = C.c_double(), C.c_double()
    lib().orc_weno5z_point(Ua, int(p), C.byref(m), C.byref(pl))
    return m.value, pl.value


def weno6ld_point(U, p=2, q=4, Cc=1.0e9, alpha_tau=35.0):
    Ua = (C.c_double * 6)(*[float(x) for x in U])
    m, pl = C.c_double(), C.c_double()
    lib().orc_weno6ld_point(Ua, int(p), int(q), C.c_double(Cc), C.c_double(alpha_tau), C.byref(m), C.byref(pl))
    return m.value, pl.value


def riemann_point(model, dim, ns, gamma, direction, V_L, V_R):
    """gamma: the species gammas; for the four-eqn conservative model followed by the species gas constants R."""
    neq = _neq_of(model, dim, ns)
    g = (C.c_double * (2 * MAX_SPECIES))(*[float(x) for x in gamma])
    VL = (C.c_double * neq)(*[float(x) for x in V_L])
    VR = (C.c_double * neq)(*[float(x) for x in V_R])
    F1 = (C.c_double * neq)()
    F2 = (C.c_double * neq)()
    vm = C.c_double()
    lib().orc_riemann_point(int(model), int(dim), int(ns), g, int(direction), VL, VR, F1, F2, C.byref(vm))
    return np.array(F1[:]), np.array(F2[:]), vm.value


def side_thermo(model, dim, ns, gamma, V):
    neq = _neq_of(model, dim, ns)
    g = (C.c_double * (2 * MAX_SPECIES))(*[float(x) for x in gamma])
    Va = (C.c_double * neq)(*[float(x) for x in V])
    r, c, e = C.c_double(), C.c_double(), C.c_double()
    lib().orc_side_thermo(int(model), int(dim), int(ns), g, Va, C.byref(r), C.byref(c), C.byref(e))
    return r.value, c.value, e.value


# ---------------------------------------------------------------------------------------------------------------------
# node-based sixth-order diffusive flux of the single-species Navier-Stokes application
# (oracle_diffusive.c)
# ---------------------------------------------------------------------------------------------------------------------
GD = 6   # DiffusiveFluxReconstructorNodeSixthOrder.cpp:24


class _Transport(C.Structure):
    _fields_ = [("mu", C.c_double), ("mu_v", C.c_double), ("c_p", C.c_double), ("c_v", C.c_double), ("Pr", C.c_double)]


@dataclass
class Transport:
    """CONSTANT shear / bulk viscosity and PRANDTL conductivity of an ideal gas (orc_transport)."""
    mu: float
    mu_v: float = 0.0
    c_p: float = 1004.5
    c_v: float = 717.5
    Pr: float = 0.72

    def c(self) -> _Transport:
        return _Transport(self.mu, self.mu_v, self.c_p, self.c_v, self.Pr)


def diff_ghost_shape(desc: PatchDesc):
    return tuple(int(desc.n[a]) + 2 * GD for a in reversed(range(desc.dim)))


def compute_diffusive_flux(desc: PatchDesc, tr: Transport, Q: np.ndarray, dt: float):
    """Q: (neq, *diff_ghost_shape) with all 6 ghost layers filled.  Returns the list per direction of (neq, *side_shape)."""
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    assert Q.shape == (desc.neq,) + diff_ghost_shape(desc), (Q.shape, diff_ghost_shape(desc))
    neq, dim = desc.neq, desc.dim
    F = [np.full((neq,) + desc.side_shape(a), np.nan) for a in range(dim)]
    d, t = desc.c(), tr.c()
    Fp = _pp([F[a][e] for a in range(dim) for e in range(neq)])
    Qp = _pp([Q[c] for c in range(neq)])
    L = lib()
    L.orc_compute_diffusive_flux.restype = C.c_int
    rc = L.orc_compute_diffusive_flux(C.byref(d), C.byref(t), Qp, C.c_double(dt), Fp)
    assert rc == 0
    return F


def compute_diffusive_flux_midpoint(desc: PatchDesc, tr: Transport, Q: np.ndarray, dt: float):
    """DiffusiveFluxReconstructorMidpointSixthOrder ("MIDPOINT_SIXTH_ORDER"); same arguments as compute_diffusive_flux (the
    cells within 5 of the interior are read)."""
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    assert Q.shape == (desc.neq,) + diff_ghost_shape(desc), (Q.shape, diff_ghost_shape(desc))
    neq, dim = desc.neq, desc.dim
    F = [np.full((neq,) + desc.side_shape(a), np.nan) for a in range(dim)]
    d, t = desc.c(), tr.c()
    L = lib()
    L.orc_compute_diffusive_flux_midpoint.restype = C.c_int
    rc = L.orc_compute_diffusive_flux_midpoint(C.byref(d), C.byref(t), _pp([Q[c] for c in range(neq)]), C.c_double(dt),
                                               _pp([F[a][e] for a in range(dim) for e in range(neq)]))
    assert rc == 0
    return F


def mid_point_kernels(u6, dx_inv, F5, dt):
    """(staggered derivative, interpolation) of six node values around a midpoint, face value of five midpoint fluxes"""
    L = lib()
    for f in (L.orc_mid_derivative, L.orc_mid_interpolate, L.orc_mid_reconstruct):
        f.restype = C.c_double
    u = (C.c_double * 6)(*[float(x) for x in u6])
    F = (C.c_double * 5)(*[float(x) for x in F5])
    return (L.orc_mid_derivative(u, C.c_double(dx_inv)), L.orc_mid_interpolate(u), L.orc_mid_reconstruct(F, C.c_double(dt)))


def mid_side_diffusivities(dim, direction, mu, mu_v, kappa, vel):
    D = (C.c_double * 8)()
    v = (C.c_double * 3)(*[float(x) for x in list(vel) + [0.0] * (3 - len(vel))])
    lib().orc_mid_side_diffusivities(int(dim), int(direction), C.c_double(mu), C.c_double(mu_v), C.c_double(kappa), v, D)
    return np.array(D[:8 if dim == 3 else 7])


def mid_side_terms(dim, fdir, ddir, e):
    n, var, diff = C.c_int(), (C.c_int * 4)(), (C.c_int * 4)()
    lib().orc_mid_side_terms(int(dim), int(fdir), int(ddir), int(e), C.byref(n), var, diff)
    return [(var[i], diff[i]) for i in range(n.value)]


def advance_stage_ns(desc: PatchDesc, g: int, alpha, beta, U_int, Fc_int, Fd_int, S_int):
    """One NavierStokes::advanceSingleStepOnPatch (conservative diffusive flux).  U_int[m]: (neq, *shape with ghost g)."""
    ncoef, neq, dim = len(alpha), desc.neq, desc.dim
    U_out = np.zeros_like(np.ascontiguousarray(U_int[0]))
    d = desc.c()
    keep = []
    PP = C.POINTER(C.POINTER(C.c_double))
    tabs = [(PP * ncoef)() for _ in range(4)]
    for m in range(ncoef):
        Um = np.ascontiguousarray(U_int[m])
        ptrs = [_pp([Um[c] for c in range(neq)])]
        for src in (Fc_int, Fd_int):
            ptrs.append(_pp([src[m][a][e] for a in range(dim) for e in range(neq)] if src[m] is not None
                            else [None] * (dim * neq)))
        ptrs.append(_pp([S_int[m][e] for e in range(neq)] if S_int[m] is not None else [None] * neq))
        keep += [Um] + ptrs
        for t, p_ in zip(tabs, ptrs):
            t[m] = C.cast(p_, PP)
    a = (C.c_double * ncoef)(*[float(x) for x in alpha])
    b = (C.c_double * ncoef)(*[float(x) for x in beta])
    L = lib()
    L.orc_advance_stage_ns.restype = C.c_int
    rc = L.orc_advance_stage_ns(C.byref(d), C.c_int(g), C.c_int(ncoef), a, b, tabs[0], tabs[1], tabs[2], tabs[3],
                                _pp([U_out[c] for c in range(neq)]))
    assert rc == 0
    return U_out


def diff_first_derivative(u7, dx_inv):
    L = lib()
    L.orc_diff_first_derivative.restype = C.c_double
    return L.orc_diff_first_derivative((C.c_double * 7)(*[float(x) for x in u7]), C.c_double(dx_inv))


def diff_reconstruct(F6, dt):
    L = lib()
    L.orc_diff_reconstruct.restype = C.c_double
    return L.orc_diff_reconstruct((C.c_double * 6)(*[float(x) for x in F6]), C.c_double(dt))


def diff_point(dim, gamma, rho, p, vel, tr: Transport):
    """(T, kappa, D[...]) of one cell."""
    L = lib()
    L.orc_diff_temperature.restype = C.c_double
    L.orc_diff_conductivity.restype = C.c_double
    T = L.orc_diff_temperature(C.c_double(gamma), C.c_double(tr.c_v), C.c_double(rho), C.c_double(p))
    kappa = L.orc_diff_conductivity(C.c_double(tr.c_p), C.c_double(tr.mu), C.c_double(tr.Pr))
    D = (C.c_double * 13)()
    L.orc_diff_diffusivities(C.c_int(dim), C.c_double(tr.mu), C.c_double(tr.mu_v), C.c_double(kappa),
                             (C.c_double * 3)(*[float(x) for x in list(vel) + [0.0] * (3 - len(vel))]), D)
    return T, kappa, list(D)[:13 if dim == 3 else 10]


def diff_terms(dim, fdir, ddir, e):
    n = C.c_int()
    var, dif = (C.c_int * 4)(), (C.c_int * 4)()
    lib().orc_diff_terms(C.c_int(dim), C.c_int(fdir), C.c_int(ddir), C.c_int(e), C.byref(n), var, dif)
    return [(var[i], dif[i]) for i in range(n.value)]


def diff_derivative_array(dim, ddir, u, n, dx_inv):
    """computeFirstDerivativesIn{X,Y,Z} over the reference's range; u on the ghost box (6); untouched entries are NaN."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.full_like(u, np.nan)
    lib().orc_diff_derivative_array(C.c_int(dim), C.c_int(ddir), u.ctypes.data_as(C.POINTER(C.c_double)),
                                    (C.c_int * 3)(*[int(x) for x in list(n) + [1] * (3 - len(n))]), C.c_double(dx_inv),
                                    out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def diff_reconstruct_array(dim, fdir, F_node, n, dt):
    """reconstructFlux{X,Y,Z} into a zero-filled side array of direction fdir."""
    F_node = np.ascontiguousarray(F_node, dtype=np.float64)
    shape = [int(x) for x in n][:dim]
    shape[fdir] += 1
    out = np.zeros(tuple(reversed(shape)))
    lib().orc_diff_reconstruct_array(C.c_int(dim), C.c_int(fdir), F_node.ctypes.data_as(C.POINTER(C.c_double)),
                                     (C.c_int * 3)(*[int(x) for x in list(n) + [1] * (3 - len(n))]), C.c_double(dt),
                                     out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def ns_spectral_radii_and_dt(desc: PatchDesc, tr: Transport, c_p_eos: float, Q: np.ndarray):
    """NavierStokes::computeSpectralRadiusesAndStableDtOnPatch on a six-ghost state: (acoustic radii per direction, stable
    dt, maximum diffusive spectral radius)."""
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    assert Q.shape == (desc.neq,) + diff_ghost_shape(desc)
    out = (C.c_double * 5)()
    d, t = desc.c(), tr.c()
    L = lib()
    L.orc_ns_spectral_radii_and_dt.restype = C.c_int
    rc = L.orc_ns_spectral_radii_and_dt(C.byref(d), C.byref(t), C.c_double(c_p_eos), _pp([Q[c] for c in range(desc.neq)]), out)
    assert rc == 0
    return list(out)[:desc.dim], out[desc.dim], out[desc.dim + 1]


def diff_max_diffusivity_point(dim, mu, mu_v, kappa, c_p_eos, rho, dx):
    L = lib()
    L.orc_diff_max_diffusivity.restype = C.c_double
    L.orc_diff_spectral_radius.restype = C.c_double
    D = L.orc_diff_max_diffusivity(C.c_double(mu), C.c_double(mu_v), C.c_double(kappa), C.c_double(c_p_eos), C.c_double(rho))
    return D, L.orc_diff_spectral_radius(C.c_int(dim), C.c_double(D), (C.c_double * 3)(*[float(x) for x in list(dx) + [1.0] * (3 - len(dx))]))
