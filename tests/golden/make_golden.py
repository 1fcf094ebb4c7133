#!/usr/bin/env python
"""Generate tests/golden/point_kernels.npz from the REFERENCE's own compiled point kernels (oracle/_ref,
built by oracle/build_ref.py from /root/reference).  Run in the build container (where /root/reference
exists); the .npz is committed so that the GPU box, which has no reference tree, can still pin the oracle.

    python tests/golden/make_golden.py

Contents (all float64, seeded):
  weno_U      (n, 6)   six-point stencils (smooth, random, shock-like, constant, tiny-variation)
  weno_minus  (n,)     performLocalWENOInterpolationMinus  (ConvectiveFluxReconstructorWCNS5-JS-HLLC-HLL.cpp:78-118)
  weno_plus   (n,)     performLocalWENOInterpolationPlus   (:124-164)
  path_points_in (n, 16), path_points_out (n, 5): first derivative (DerivativeFirstOrder.cpp:601), dilatation, vorticity
                       magnitude, sensor value and face flux (ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:1631,
                       1653-1657, 2098-2101, 2370-2375) from the reference's own statements, compiled verbatim
  path_points2_in (n, 32), path_points2_out (n, 20): velocity and specific internal energy (FlowModelSingleSpecies.cpp:
                       2824-2826, 3049-3051), face averages, characteristic projection and back-projection
                       (FlowModelBasicUtilitiesSingleSpecies.cpp:5000-5001, 6324-6329, 7373-7379), RK alpha/beta update
                       (Euler.cpp:1479, 1544-1548), again the reference's own statements compiled verbatim
  path_points3_in (n, 56), path_points3_out (n, 32): five-eqn Allaire, two species, 3-D, x: mixture density, mass
                       fractions, velocity, internal energy, mixture gamma, pressure, sound speed, face averages,
                       projection / back-projection, advective source (EquationOfStateMixingRules.cpp:871,
                       FlowModelFiveEqnAllaire.cpp:3965, 4188-4190, 4428-4430, 4700-4853,
                       EquationOfStateMixingRulesIdealGas.cpp:7544-7586, EquationOfStateIdealGas.cpp:5756, 8157, 8308,
                       FlowModelBasicUtilitiesFiveEqnAllaire.cpp:7952-7996, 8848-8921, 9700-9760,
                       ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:2623-2641), statements compiled verbatim
  path_points4_in (n, 12), path_points4_out (n, 4): (rho, c, epsilon) the five-eqn Riemann solver rebuilds from one
                       interpolated side (FlowModelRiemannSolverFiveEqnAllaireHLLC.cpp:5709-5941, mixture gamma from the
                       ns-1 volume fractions EquationOfStateMixingRulesIdealGas.cpp:7770-7829, EquationOfStateIdealGas.cpp:
                       8157, 8308, 6414), statements compiled verbatim
  path_points5_in (n, 16), path_points5_out (n, 2): bounds flags of one interpolated side, five-eqn (3-D x / y / z blocks,
                       FlowModelBasicUtilitiesFiveEqnAllaire.cpp:6400-7340) and single-species
                       (FlowModelBasicUtilitiesSingleSpecies.cpp:3311-3326): the reference's if / else blocks verbatim
  path_points6_in (n, 8), path_points6_out (n, 6): max wave speeds, spectral radii, their sum, running maximum and stable
                       dt of one cell (FlowModelSingleSpecies.cpp:4064, 4237, 4365; Euler.cpp:846-861), statements verbatim
  ref_constants (7,): HAMERS_EPSILON, sensor threshold, Y bounds lo/up, Z bounds lo/up, ghost width, parsed from the source
  rk_alpha, rk_beta, rk_gamma (3, 3): the reference's default SSPRK(3,3) table (RungeKuttaLevelIntegrator.cpp:3894-3929)
  eos_in (n, 3) = (gamma, rho, epsilon), eos_out (n, 3) = (p, c, epsilon from p): EquationOfStateIdealGas scalar members
  rp_<case>_{VL,VR,thermo,F_HLLC,F_HYB,vel_mid} for case in ss2d{0,1}, ss3d{0,1,2}, fe2d{0,1}, fe3d{0,1,2}:
                       computeLocal...HLLC{2D,3D} / ...HLLC_HLL{2D,3D} of the reference
                       (FlowModelRiemannSolverSingleSpeciesHLLC.cpp:604-1079, ...HLLC-HLL.cpp:889-1629,
                        FlowModelRiemannSolverFiveEqnAllaireHLLC.cpp:712-1591, ...HLLC-HLL.cpp:1283-2440)
                       thermo = (rho_L, rho_R, c_L, c_R, eps_L, eps_R) as fed to the reference kernels
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libhamers_ref.so")


def ref_lib():
    lib = C.CDLL(REF_SO)
    lib.ref_riemann_point.restype = C.c_int
    return lib


def weno_inputs(rng, n=400):
    U = []
    for _ in range(n // 5):
        x0 = rng.uniform(-1, 1)
        h = 10.0 ** rng.uniform(-3, -0.5)
        U.append(np.sin(np.pi * (x0 + h * np.arange(6))) + 1.5)                  # smooth
        U.append(rng.uniform(0.1, 3.0, 6))                                       # rough
        s = rng.uniform(0.5, 2.0, 6)
        s[rng.integers(1, 5):] *= rng.uniform(3.0, 20.0)                         # jump inside the stencil
        U.append(s)
        U.append(np.full(6, rng.uniform(-2, 2)))                                 # constant
        U.append(rng.uniform(0.5, 2.0) * (1.0 + 1e-9 * rng.standard_normal(6)))  # round-off level variation
    return np.array(U)


def ref_weno(lib, U, p=2):
    m, pl = C.c_double(), C.c_double()
    lib.ref_weno5js_point((C.c_double * 6)(*U), int(p), C.byref(m), C.byref(pl))
    return m.value, pl.value


def ref_weno_z(lib, U, p=2):
    m, pl = C.c_double(), C.c_double()
    lib.ref_weno5z_point((C.c_double * 6)(*U), int(p), C.byref(m), C.byref(pl))
    return m.value, pl.value


def ref_weno_ld(lib, U, p=2, q=4, Cc=1.0e9, alpha_tau=35.0):
    m, pl = C.c_double(), C.c_double()
    lib.ref_weno6ld_point((C.c_double * 6)(*U), int(p), int(q), C.c_double(Cc), C.c_double(alpha_tau), C.byref(m), C.byref(pl))
    return m.value, pl.value


def riemann_inputs(rng, model, dim, ns, n=120):
    neq = dim + 2 if model == 0 else dim + 2 * ns
    VL, VR = np.zeros((n, neq)), np.zeros((n, neq))
    for V in (VL, VR):
        nm = 1 if model == 0 else ns
        V[:, :nm] = rng.uniform(0.2, 3.0, (n, nm))
        V[:, nm:nm + dim] = rng.uniform(-3.0, 3.0, (n, dim))
        V[:, nm + dim] = rng.uniform(0.2, 10.0, n)
        if model == 1:
            V[:, nm + dim + 1:] = rng.uniform(0.05, 0.95, (n, ns - 1))
    # identical velocities (|du| < eps branch), supersonic both ways (upwind overrides), strong compression
    VR[0, :] = VL[0, :]
    iv = 1 if model == 0 else ns
    VL[1, iv:iv + dim] = 8.0
    VR[1, iv:iv + dim] = 8.5
    VL[2, iv:iv + dim] = -8.0
    VR[2, iv:iv + dim] = -8.5
    VL[3, iv:iv + dim] = 2.5
    VR[3, iv:iv + dim] = -2.5
    return VL, VR


def ref_riemann(lib, model, dim, ns, gamma, direction, VL, VR):
    neq = VL.shape[0]
    rL, cL, eL = orc.side_thermo(model, dim, ns, gamma, VL)
    rR, cR, eR = orc.side_thermo(model, dim, ns, gamma, VR)
    F1, F2, vm = (C.c_double * neq)(), (C.c_double * neq)(), C.c_double()
    rc = lib.ref_riemann_point(model, dim, ns, direction, (C.c_double * neq)(*VL), (C.c_double * neq)(*VR),
                               C.c_double(rL), C.c_double(rR), C.c_double(cL), C.c_double(cR),
                               C.c_double(eL), C.c_double(eR), F1, F2, C.byref(vm))
    assert rc == 0
    return np.array(F1[:]), np.array(F2[:]), vm.value, (rL, rR, cL, cR, eL, eR)


CASES = [("ss", 0, 2, 1, (1.4,)), ("ss", 0, 3, 1, (1.4,)), ("fe", 1, 2, 2, (1.6, 1.4)), ("fe", 1, 3, 2, (1.6, 1.4))]


def main():
    if not os.path.exists(REF_SO):
        raise SystemExit("oracle/_ref/libhamers_ref.so missing: run oracle/build_ref.py where /root/reference exists")
    orc.build()
    lib = ref_lib()
    rng = np.random.default_rng(20261017)
    out = {}
    U = weno_inputs(rng)
    res = np.array([ref_weno(lib, u) for u in U])
    out["weno_U"], out["weno_minus"], out["weno_plus"] = U, res[:, 0], res[:, 1]
    res3 = np.array([ref_weno(lib, u, 3) for u in U])
    out["weno_minus_p3"], out["weno_plus_p3"] = res3[:, 0], res3[:, 1]
    # WCNS5-Z (ConvectiveFluxReconstructorWCNS5-Z-HLLC-HLL.cpp:78-165) and WCNS6-LD
    # (ConvectiveFluxReconstructorWCNS6-LD-HLLC-HLL.cpp:143-339) on the same stencils; LD also with a low alpha_tau so
    # that the R_tau > alpha_tau blend is taken on the smooth stencils too
    rz = np.array([ref_weno_z(lib, u) for u in U])
    out["weno_z_minus"], out["weno_z_plus"] = rz[:, 0], rz[:, 1]
    for tag, args in (("ld", (2, 4, 1.0e9, 35.0)), ("ld_b", (2, 4, 1.0e3, 0.5)), ("ld_c", (3, 2, 10.0, 2.0))):
        rl = np.array([ref_weno_ld(lib, u, *args) for u in U])
        out[f"weno_{tag}_minus"], out[f"weno_{tag}_plus"] = rl[:, 0], rl[:, 1]
        out[f"weno_{tag}_params"] = np.array(args, dtype=np.float64)
    # the default Runge-Kutta table, SSPRK(3,3), read from the reference's source text (RungeKuttaLevelIntegrator.cpp:3894-3929)
    import re
    ref_root = os.environ.get("HAMERS_REFERENCE", "/root/reference")
    with open(os.path.join(ref_root, "src", "algs", "integrator", "RungeKuttaLevelIntegrator.cpp")) as fh:
        txt = fh.read()
    blk = txt[txt.index("Use SSPRK(3, 3) Runge-Kutta scheme as the default scheme"):]
    blk = blk[:blk.index("else if (input_db)")]
    for nm in ("alpha", "beta", "gamma"):
        tab = np.zeros((3, 3))
        for i, j, expr in re.findall(r"d_" + nm + r"\[(\d)\]\[(\d)\]\s*=\s*([0-9./ ]+);", blk):
            num = [float(x) for x in expr.split("/")]
            tab[int(i), int(j)] = num[0] / num[1] if len(num) == 2 else num[0]
        out["rk_" + nm] = tab
    # constants of the hard switches, read from the reference's source text
    def grab(rel, pattern):
        with open(os.path.join(ref_root, rel)) as fh2:
            return float(re.search(pattern, fh2.read()).group(1))
    wcns56 = "src/flow/convective_flux_reconstructors/WCNS56/ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp"
    fe_hpp = "include/flow/flow_models/five-eqn_Allaire/FlowModelBasicUtilitiesFiveEqnAllaire.hpp"
    out["ref_constants"] = np.array([
        grab("include/HAMeRS_config.hpp.in", r"#define HAMERS_EPSILON\s+([0-9.eE+-]+)"),
        grab(wcns56, r"s_z\[idx_midpoint_z\] > ([0-9.]+)"),
        grab(fe_hpp, r"d_Y_bound_lo = double\(([-0-9.]+)\)"), grab(fe_hpp, r"d_Y_bound_up = double\(([-0-9.]+)\)"),
        grab(fe_hpp, r"d_Z_bound_lo = double\(([-0-9.]+)\)"), grab(fe_hpp, r"d_Z_bound_up = double\(([-0-9.]+)\)"),
        grab(wcns56, r"d_num_conv_ghosts = hier::IntVector::getOne\(d_dim\)\*([0-9]+)")])
    # the reference's own statements of the sensor chain and of the face-flux formula (oracle/build_ref.py: path_statements)
    rng_pp = np.random.default_rng(123)
    pp_in = rng_pp.standard_normal((400, 16)) * 10.0 ** rng_pp.uniform(-3, 3, (400, 1))
    pp_in[:, 2] = np.abs(pp_in[:, 2]) + 1.0e-3       # dx
    pp_in[:, 12] = np.abs(pp_in[:, 12])              # dt
    pp_out = []
    for v in pp_in:
        o = (C.c_double * 5)()
        lib.ref_path_points((C.c_double * 16)(*v), o)
        pp_out.append(list(o))
    out["path_points_in"], out["path_points_out"] = pp_in, np.array(pp_out)
    # second group (oracle/build_ref.py: path_statements2); own generator again
    rng_p2 = np.random.default_rng(321)
    p2_in = rng_p2.standard_normal((400, 32)) * 10.0 ** rng_p2.uniform(-2, 2, (400, 1))
    pos = [0, 5, 6, 7, 8, 29, 30, 31]               # densities, sound speeds, mesh widths
    p2_in[:, pos] = np.abs(p2_in[:, pos]) + 1.0e-3
    p2_out = []
    for v in p2_in:
        o = (C.c_double * 20)()
        lib.ref_path_points2((C.c_double * 32)(*v), o)
        p2_out.append(list(o))
    out["path_points2_in"], out["path_points2_out"] = p2_in, np.array(p2_out)
    # third group, five-eqn with two species (oracle/build_ref.py: path_statements3); own generator
    rng_p3 = np.random.default_rng(654)
    p3_in = rng_p3.standard_normal((400, 56)) * 10.0 ** rng_p3.uniform(-1, 1, (400, 1))
    pos3 = [0, 1, 5, 6, 7, 10, 11, 12, 13, 14, 15, 16, 17, 33, 53, 54, 55]     # partial densities, E, Z, averages' inputs, dt, dx
    p3_in[:, pos3] = np.abs(p3_in[:, pos3]) + 1.0e-3
    p3_in[:, 8:10] = rng_p3.uniform(1.1, 1.7, (400, 2))                      # species gammas
    p3_in[:, 5] += 0.5 * (p3_in[:, 2:5] ** 2).sum(axis=1) / p3_in[:, 0:2].sum(axis=1)   # positive internal energy
    p3_out = []
    for v in p3_in:
        o = (C.c_double * 32)()
        lib.ref_path_points3((C.c_double * 56)(*v), o)
        p3_out.append(list(o))
    out["path_points3_in"], out["path_points3_out"] = p3_in, np.array(p3_out)
    assert np.isfinite(out["path_points3_out"]).all()
    # fourth group: five-eqn thermodynamic state of one interpolated side (oracle/build_ref.py: path_statements4)
    rng_p4 = np.random.default_rng(987)
    p4_in = np.zeros((400, 12))
    p4_in[:, 0:2] = 10.0 ** rng_p4.uniform(-3, 1, (400, 2))            # partial densities
    p4_in[:, 2:5] = rng_p4.standard_normal((400, 3))                   # velocity
    p4_in[:, 5] = 10.0 ** rng_p4.uniform(-2, 2, 400)                   # pressure
    p4_in[:, 6] = rng_p4.uniform(-0.001, 1.001, 400)                   # interpolated volume fraction, within the bounds
    p4_in[:, 7:9] = rng_p4.uniform(1.1, 1.7, (400, 2))                 # species gammas
    p4_out = []
    for v in p4_in:
        o = (C.c_double * 4)()
        lib.ref_path_points4((C.c_double * 12)(*v), o)
        p4_out.append(list(o))
    out["path_points4_in"], out["path_points4_out"] = p4_in, np.array(p4_out)
    assert np.isfinite(out["path_points4_out"]).all()
    # fifth group: bounds flags of one interpolated side (oracle/build_ref.py: path_statements5), all outcomes and the
    # directions in which the reference's five-eqn c^2 check differs
    rng_p5 = np.random.default_rng(555)
    p5_in = np.zeros((1200, 16))
    for v in p5_in:
        v[0:2] = rng_p5.uniform(-0.05, 1.0, 2)
        v[2:5] = rng_p5.standard_normal(3)
        v[5] = rng_p5.uniform(-0.1, 2.0)
        v[6] = [rng_p5.uniform(-0.3, 1.3), rng_p5.uniform(-8.0, 8.0), rng_p5.uniform(-1500.0, 1500.0)][rng_p5.integers(0, 3)]
        v[7:9] = [rng_p5.uniform(1.1, 1.7, 2), np.array([1.4, 1.09]), np.array([1.0005, 3.0])][rng_p5.integers(0, 3)]
        v[9] = rng_p5.integers(0, 3)
        v[10] = rng_p5.uniform(-0.1, 1.0)
        v[11:14] = rng_p5.standard_normal(3)
        v[14] = rng_p5.uniform(-0.1, 1.0)
    p5_in[::50, 0] = 0.0                                  # the comparisons are strict
    p5_in[::77, 14] = 0.0
    p5_in[::91, 5] = 0.0
    p5_in[::97, 6] = [-1000.0, 1000.0, -0.0, 1.0][0]
    p5_out = []
    for v in p5_in:
        o = (C.c_double * 2)()
        lib.ref_path_points5((C.c_double * 16)(*v), o)
        p5_out.append(list(o))
    out["path_points5_in"], out["path_points5_out"] = p5_in, np.array(p5_out)
    # sixth group: spectral radii and stable dt of one cell (oracle/build_ref.py: path_statements6)
    rng_p6 = np.random.default_rng(666)
    p6_in = np.abs(rng_p6.standard_normal((300, 8))) * 10.0 ** rng_p6.uniform(-2, 2, (300, 1)) + 1.0e-3
    p6_in[:, 0:3] *= rng_p6.choice([-1.0, 1.0], (300, 3))
    p6_out = []
    for v in p6_in:
        o = (C.c_double * 6)()
        lib.ref_path_points6((C.c_double * 8)(*v), o)
        p6_out.append(list(o))
    out["path_points6_in"], out["path_points6_out"] = p6_in, np.array(p6_out)
    # ideal-gas EOS scalars (EquationOfStateIdealGas.cpp:29-45, 561-577, 1093-1108); own generator: the arrays above keep
    # their values
    rng_eos = np.random.default_rng(77)
    eos_in = np.stack([rng_eos.uniform(1.1, 1.7, 300), 10.0 ** rng_eos.uniform(-3, 3, 300), 10.0 ** rng_eos.uniform(-3, 4, 300)], axis=1)
    eos_out = []
    for g, r, e in eos_in:
        p_, c_, b_ = C.c_double(), C.c_double(), C.c_double()
        lib.ref_eos_point(C.c_double(g), C.c_double(r), C.c_double(e), C.byref(p_), C.byref(c_), C.byref(b_))
        eos_out.append((p_.value, c_.value, b_.value))
    out["eos_in"], out["eos_out"] = eos_in, np.array(eos_out)
    for tag, model, dim, ns, gam in CASES:
        for d in range(dim):
            VL, VR = riemann_inputs(rng, model, dim, ns)
            F1s, F2s, vms, ths = [], [], [], []
            for a, b in zip(VL, VR):
                F1, F2, vm, th = ref_riemann(lib, model, dim, ns, gam, d, a, b)
                F1s.append(F1), F2s.append(F2), vms.append(vm), ths.append(th)
            key = f"rp_{tag}{dim}d{d}"
            out[key + "_VL"], out[key + "_VR"] = VL, VR
            out[key + "_thermo"] = np.array(ths)
            out[key + "_F_HLLC"], out[key + "_F_HYB"], out[key + "_vel_mid"] = np.array(F1s), np.array(F2s), np.array(vms)
    path = os.path.join(HERE, "point_kernels.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("weno")})


if __name__ == "__main__":
    main()
