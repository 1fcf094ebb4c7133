"""ctypes front end of the TEST-ONLY host emulation of the CUDA sweeps (tests/host_emu/emu.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "host_emu", "libhb2_emu.so")          # scheme 0; libhb2_emu_s<k>.so for WCNS5-Z / WCNS6-LD
_SRC = os.path.join(_HERE, "host_emu", "emu.cpp")
_CORE = [os.path.join(_HERE, "..", "hamers_b200", "csrc", f) for f in ("hb2_core.cuh", "hb2_fast.cuh", "hb2_sweep.cuh", "hb2_sensor.cuh")]
_LIB = {}


class EmuDesc(C.Structure):
    _fields_ = [("dim", C.c_int), ("n", C.c_int * 3), ("model", C.c_int), ("ns", C.c_int),
                ("gamma", C.c_double * 4), ("dx", C.c_double * 3), ("weno_p", C.c_int),
                ("math", C.c_int), ("bx", C.c_int), ("seg_len", C.c_int),
                ("weno_q", C.c_int), ("weno_C", C.c_double), ("weno_alpha_tau", C.c_double), ("ghosts", C.c_int),
                ("R", C.c_double * 4)]


def lib(scheme: int = 0):
    """The emulation library of one nonlinear interpolator (HB2_SCHEME is a compile-time property of the kernels)."""
    if scheme not in _LIB:
        so = _SO if scheme == 0 else _SO.replace(".so", f"_s{scheme}.so")
        stale = (not os.path.exists(so)) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in [_SRC] + _CORE)
        if stale:
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                                   "-Wno-unknown-pragmas", f"-DHB2_SCHEME={scheme}", "-o", so, _SRC])
        _LIB[scheme] = C.CDLL(so)
    return _LIB[scheme]


def _pp(arrs):
    P = (C.POINTER(C.c_double) * len(arrs))()
    for i, a in enumerate(arrs):
        if a is not None:
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
            P[i] = a.ctypes.data_as(C.POINTER(C.c_double))
    return P


def _desc(desc, math, bx, seg_len, ghosts=0):
    d = EmuDesc()
    d.ghosts = ghosts
    d.dim = desc.dim
    for a in range(3):
        d.n[a] = int(desc.n[a]) if a < desc.dim else 1
        d.dx[a] = float(desc.dx[a]) if a < desc.dim else 1.0
    d.model, d.ns = desc.model, desc.ns
    for i, g in enumerate(desc.gamma):
        d.gamma[i] = g
    for i, r in enumerate(getattr(desc, "R", ())):
        d.R[i] = r
    d.weno_p, d.math, d.bx, d.seg_len = desc.weno_p, math, bx, seg_len
    d.weno_q, d.weno_C, d.weno_alpha_tau = desc.weno_q, desc.weno_C, desc.weno_alpha_tau
    return d


def flux_and_source(desc, Q, dt, math=0, bx=0, seg_len=0, source=None, ghosts=0):
    """ghosts: ghost width of the layout of Q (0 = the convective 4)."""
    neq, dim = desc.neq, desc.dim
    F = [np.full((neq,) + desc.side_shape(a), np.nan) for a in range(dim)]
    S = np.zeros((neq,) + desc.cell_shape) if source is None else source
    d = _desc(desc, math, bx, seg_len, ghosts)
    Q = np.ascontiguousarray(Q)
    rc = lib(desc.scheme).emu_flux_and_source(C.byref(d), _pp([Q[c] for c in range(desc.ncomp)]), C.c_double(dt),
                                   _pp([F[a][e] for a in range(dim) for e in range(neq)]),
                                   _pp([S[e] for e in range(neq)]))
    assert rc == 0
    return F, S


def fused_stage(desc, alpha, beta, U_int, dt, math=0, bx=0, seg_len=0, push=False, ghosts=0):
    ncoef = len(alpha)
    U_out = np.zeros_like(np.ascontiguousarray(U_int[0]))
    d = _desc(desc, math, bx, seg_len, ghosts)
    Us = [np.ascontiguousarray(u) for u in U_int]
    tab = _pp([Us[m][c] for m in range(ncoef) for c in range(desc.ncomp)])
    a = (C.c_double * ncoef)(*[float(x) for x in alpha])
    b = (C.c_double * ncoef)(*[float(x) for x in beta])
    rc = lib(desc.scheme).emu_fused_stage_push(C.byref(d), ncoef, a, b, tab, C.c_double(dt), _pp([U_out[c] for c in range(desc.ncomp)]),
                                    1 if push else 0)
    assert rc == 0
    return U_out


# ---- diffusive-flux kernels (tests/host_emu/emu_diffusive.cpp) ----------------------------------------
_DSO = os.path.join(_HERE, "host_emu", "libhb2_emu_diffusive.so")
_DSRC = os.path.join(_HERE, "host_emu", "emu_diffusive.cpp")
_DCORE = os.path.join(_HERE, "..", "hamers_b200", "csrc", "hb2_diffusive.cuh")
_DLIB = None


class EmuDiffDesc(C.Structure):
    _fields_ = [("dim", C.c_int), ("n", C.c_int * 3), ("dx", C.c_double * 3), ("gamma", C.c_double), ("c_v", C.c_double),
                ("mu", C.c_double), ("mu_v", C.c_double), ("c_p", C.c_double), ("Pr", C.c_double)]


def dlib():
    global _DLIB
    if _DLIB is None:
        stale = (not os.path.exists(_DSO)) or any(os.path.getmtime(s) > os.path.getmtime(_DSO) for s in (_DSRC, _DCORE))
        if stale:
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                   "-o", _DSO, _DSRC])
        _DLIB = C.CDLL(_DSO)
    return _DLIB


def _ddesc(desc, tr):
    d = EmuDiffDesc()
    d.dim = desc.dim
    for a in range(3):
        d.n[a] = int(desc.n[a]) if a < desc.dim else 1
        d.dx[a] = float(desc.dx[a]) if a < desc.dim else 1.0
    d.gamma, d.c_v, d.mu, d.mu_v, d.c_p, d.Pr = desc.gamma[0], tr.c_v, tr.mu, tr.mu_v, tr.c_p, tr.Pr
    return d


def diffusive_flux(desc, tr, Q, dt):
    neq, dim = desc.neq, desc.dim
    F = [np.full((neq,) + desc.side_shape(a), np.nan) for a in range(dim)]
    Q = np.ascontiguousarray(Q)
    d = _ddesc(desc, tr)
    rc = dlib().emu_diffusive_flux(C.byref(d), _pp([Q[c] for c in range(neq)]), C.c_double(dt),
                                   _pp([F[a][e] for a in range(dim) for e in range(neq)]))
    assert rc == 0
    return F


def diffusive_flux_midpoint(desc, tr, Q, dt):
    """the midpoint family (DiffusiveFluxReconstructorMidpointSixthOrder) through the product's thread functions"""
    neq, dim = desc.neq, desc.dim
    F = [np.full((neq,) + desc.side_shape(a), np.nan) for a in range(dim)]
    Q = np.ascontiguousarray(Q)
    d = _ddesc(desc, tr)
    rc = dlib().emu_diffusive_flux_midpoint(C.byref(d), _pp([Q[c] for c in range(neq)]), C.c_double(dt),
                                            _pp([F[a][e] for a in range(dim) for e in range(neq)]))
    assert rc == 0
    return F


def advance_stage_ns(desc, tr, g, alpha, beta, U_int, Fc_int, Fd_int, S_int):
    ncoef, neq, dim = len(alpha), desc.neq, desc.dim
    U_out = np.full_like(np.ascontiguousarray(U_int[0]), np.nan)
    U_int = [np.ascontiguousarray(u) for u in U_int]
    d = _ddesc(desc, tr)
    rc = dlib().emu_advance_stage_ns(
        C.byref(d), C.c_int(g), C.c_int(ncoef), (C.c_double * ncoef)(*alpha), (C.c_double * ncoef)(*beta),
        _pp([U_int[m][e] for m in range(ncoef) for e in range(neq)]),
        _pp([Fc_int[m][a][e] if Fc_int[m] is not None else None for m in range(ncoef) for a in range(dim) for e in range(neq)]),
        _pp([Fd_int[m][a][e] if Fd_int[m] is not None else None for m in range(ncoef) for a in range(dim) for e in range(neq)]),
        _pp([S_int[m][e] if S_int[m] is not None else None for m in range(ncoef) for e in range(neq)]),
        _pp([U_out[e] for e in range(neq)]))
    assert rc == 0
    return U_out


def diff_fill_periodic(desc, tr, U, mask=7):
    """in place on U (neq, *six-ghost shape)"""
    d = _ddesc(desc, tr)
    assert U.flags["C_CONTIGUOUS"]
    rc = dlib().emu_diff_fill_periodic(C.byref(d), _pp([U[c] for c in range(desc.neq)]), C.c_int(mask))
    assert rc == 0
    return U


def diff_extract_view(desc, tr, U, g):
    d = _ddesc(desc, tr)
    U = np.ascontiguousarray(U)
    V = np.full((desc.neq,) + tuple(int(desc.n[a]) + 2 * g for a in reversed(range(desc.dim))), np.nan)
    rc = dlib().emu_diff_extract_view(C.byref(d), _pp([U[c] for c in range(desc.neq)]), C.c_int(g),
                                      _pp([V[c] for c in range(desc.neq)]))
    assert rc == 0
    return V


def diff_accumulate(desc, tr, g, beta, Fd, U):
    """in place on U (neq, *shape with g ghosts)"""
    d = _ddesc(desc, tr)
    assert U.flags["C_CONTIGUOUS"]
    rc = dlib().emu_diff_accumulate(C.byref(d), C.c_int(g), C.c_double(beta),
                                    _pp([Fd[a][e] for a in range(desc.dim) for e in range(desc.neq)]),
                                    _pp([U[e] for e in range(desc.neq)]))
    assert rc == 0
    return U


def diff_terms(dim, f, d, e):
    var, dif = (C.c_int * 4)(), (C.c_int * 4)()
    n = dlib().emu_diff_terms(dim, f, d, e, var, dif)
    return [(var[i], dif[i]) for i in range(n)]


def diff_divergence_accumulate(desc, tr, Q, dt, g, beta, U):
    """in place on U: U += beta (-div F_d(Q)), no side flux written"""
    d = _ddesc(desc, tr)
    Q = np.ascontiguousarray(Q)
    assert U.flags["C_CONTIGUOUS"]
    rc = dlib().emu_diff_divergence_accumulate(C.byref(d), _pp([Q[c] for c in range(desc.neq)]), C.c_double(dt), C.c_int(g),
                                               C.c_double(beta), _pp([U[e] for e in range(desc.neq)]))
    assert rc == 0
    return U


def diff_divergence_accumulate_fast(desc, tr, Q, dt, g, beta, U):
    """the same update in the re-associated arithmetic of the HB2_MATH_FAST route (3-D)"""
    d = _ddesc(desc, tr)
    Q = np.ascontiguousarray(Q)
    assert U.flags["C_CONTIGUOUS"]
    rc = dlib().emu_diff_divergence_accumulate_fast(C.byref(d), _pp([Q[c] for c in range(desc.neq)]), C.c_double(dt), C.c_int(g),
                                                    C.c_double(beta), _pp([U[e] for e in range(desc.neq)]))
    assert rc == 0
    return U


def diff_max_spectral_radius(desc, tr, c_p_eos, Q):
    d = _ddesc(desc, tr)
    f = dlib().emu_diff_max_spectral_radius
    f.restype = C.c_double
    rho = np.ascontiguousarray(Q[0])
    return f(C.byref(d), C.c_double(c_p_eos), rho.ctypes.data_as(C.POINTER(C.c_double)))


# ---- AMR operator kernels (tests/host_emu/emu_amr.cpp) --------------------------------------------------------
_ASO = os.path.join(_HERE, "host_emu", "libhb2_emu_amr.so")
_ASRC = os.path.join(_HERE, "host_emu", "emu_amr.cpp")
_ACORE = os.path.join(_HERE, "..", "hamers_b200", "csrc", "hb2_amr.cuh")
_ALIB = None


class EmuPair(C.Structure):
    _fields_ = [("dim", C.c_int), ("nc", C.c_int * 3), ("nf", C.c_int * 3), ("ratio", C.c_int * 3), ("origin", C.c_int * 3),
                ("ghosts_c", C.c_int), ("ghosts_f", C.c_int), ("ncomp", C.c_int), ("neq", C.c_int),
                ("dxc", C.c_double * 3), ("dxf", C.c_double * 3)]


def alib():
    global _ALIB
    if _ALIB is None:
        stale = (not os.path.exists(_ASO)) or any(os.path.getmtime(s) > os.path.getmtime(_ASO) for s in (_ASRC, _ACORE))
        if stale:
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                   "-o", _ASO, _ASRC])
        _ALIB = C.CDLL(_ASO)
    return _ALIB


def amr_pair(dim, nc, nf, ratio, origin, dxc, dxf, ncomp, neq, g=4):
    p = EmuPair()
    p.dim = dim
    for a in range(dim):
        p.nc[a], p.nf[a], p.ratio[a], p.origin[a], p.dxc[a], p.dxf[a] = nc[a], nf[a], ratio[a], origin[a], dxc[a], dxf[a]
    p.ghosts_c = p.ghosts_f = g
    p.ncomp, p.neq = ncomp, neq
    return p


def _i3(v, dim, fill):
    return (C.c_int * 3)(*[int(v[a]) if a < dim else fill for a in range(3)])


def amr_refine(p, Uold, Unew, tfrac, lo, hi, Uf, reverse=False):
    new = _pp([Unew[c] for c in range(p.ncomp)]) if Unew is not None else None
    rc = alib().emu_amr_refine(C.byref(p), _pp([Uold[c] for c in range(p.ncomp)]), new, C.c_double(tfrac), _i3(lo, p.dim, 0),
                               _i3(hi, p.dim, 1), _pp([Uf[c] for c in range(p.ncomp)]), 1 if reverse else 0)
    assert rc == 0


def amr_coarsen(p, Uf, lo, hi, Uc):
    rc = alib().emu_amr_coarsen(C.byref(p), _pp([Uf[c] for c in range(p.ncomp)]), _i3(lo, p.dim, 0), _i3(hi, p.dim, 1),
                                _pp([Uc[c] for c in range(p.ncomp)]))
    assert rc == 0


def amr_fluxsum(p, F, fsum):
    """F: list per direction of (neq, ...) side arrays; fsum: list [2 dir + side] of (neq, ...) arrays, updated in place."""
    rc = alib().emu_amr_fluxsum(C.byref(p), _pp([F[d][e] for d in range(p.dim) for e in range(p.neq)]),
                                _pp([fsum[k][e] for k in range(2 * p.dim) for e in range(p.neq)]))
    assert rc == 0


def amr_coarsen_fluxsum(p, fsum, Fc):
    rc = alib().emu_amr_coarsen_fluxsum(C.byref(p), _pp([fsum[k][e] for k in range(2 * p.dim) for e in range(p.neq)]),
                                        _pp([Fc[d][e] for d in range(p.dim) for e in range(p.neq)]))
    assert rc == 0


def amr_extrapolate(dim, n, g, U, direction, side):
    rc = alib().emu_amr_extrapolate(dim, _i3(n, dim, 1), g, U.shape[0], _pp([U[c] for c in range(U.shape[0])]), direction, side)
    assert rc == 0
