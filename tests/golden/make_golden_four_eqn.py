#!/usr/bin/env python
"""Generate tests/golden/four_eqn_kernels.npz from oracle/_ref/libhamers_ref.so, i.e. from the reference's
OWN code for the FOUR_EQN_CONSERVATIVE flow model compiled by oracle/build_ref.py (needs /root/reference; run in the build
container):

    python tests/golden/make_golden_four_eqn.py

Contents (inputs and the reference's outputs):
  pp7_in / pp7_out   mixture chain of a cell (rho, Y, epsilon, c_p, c_v, gamma, p, Psi_i, c), the (rho, c, epsilon) of an
                     interpolated side and its bounds flag (FlowModelFourEqnConservative.cpp, EquationOfStateMixingRulesIdealGas.cpp,
                     FlowModelBasicUtilitiesFourEqnConservative.cpp:4380-4510), two species, 3-D
  pp8_in / pp8_out   face averages, characteristic projection and back-projection, 3-D x
                     (FlowModelBasicUtilitiesFourEqnConservative.cpp:4638-6826)
  rp_fc{dim}d{dir}_* the HLLC / HLLC-HLL point kernels (FlowModelRiemannSolverFourEqnConservativeHLLC.cpp,
                     ...HLLC-HLL.cpp) fed with the oracle's side thermodynamics (pinned by pp7)
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
GAMMA_R = (1.6, 1.4, 1.2, 0.7)      # species gammas, species gas constants


def pp7_inputs(rng, n=300):
    v = np.zeros((n, 16))
    v[:, 0:2] = rng.uniform(0.05, 2.0, (n, 2))
    v[:, 2:5] = rng.uniform(-2.0, 2.0, (n, 3))
    v[:, 5] = rng.uniform(4.0, 12.0, n)
    v[:, 6:10] = GAMMA_R
    v[: n // 3, 6:8] = rng.uniform(1.1, 1.9, (n // 3, 2))
    v[: n // 3, 8:10] = rng.uniform(0.3, 3.0, (n // 3, 2))
    # interpolated side: partial densities around zero, pressure around zero -> both outcomes of every comparison
    v[:, 10:12] = rng.uniform(-0.05, 2.0, (n, 2))
    v[:, 12:15] = rng.uniform(-2.0, 2.0, (n, 3))
    v[:, 15] = rng.uniform(-0.2, 5.0, n)
    k = n // 6
    v[:k, 10] = rng.uniform(-0.003, 0.003, k) * v[:k, 11]          # Y_0 near its lower bound
    v[k:2 * k, 11] = -rng.uniform(0.0, 0.002, k) * v[k:2 * k, 10]  # Y_1 slightly negative
    return v


def pp8_inputs(rng, n=200):
    v = np.zeros((n, 24))
    v[:, 0:4] = rng.uniform(0.05, 2.0, (n, 4))
    v[:, 4:6] = rng.uniform(0.3, 3.0, (n, 2))
    v[:, 6:8] = rng.uniform(0.5, 3.0, (n, 2))
    v[:, 8:10] = rng.uniform(0.05, 2.0, (n, 2))
    v[:, 10:13] = rng.uniform(-2.0, 2.0, (n, 3))
    v[:, 13] = rng.uniform(0.2, 8.0, n)
    v[:, 14:20] = rng.uniform(-3.0, 3.0, (n, 6))
    return v


def riemann_inputs(rng, dim, ns, n=120):
    neq = dim + 1 + ns
    VL, VR = np.zeros((n, neq)), np.zeros((n, neq))
    for V in (VL, VR):
        V[:, :ns] = rng.uniform(0.1, 3.0, (n, ns))
        V[:, ns:ns + dim] = rng.uniform(-3.0, 3.0, (n, dim))
        V[:, ns + dim] = rng.uniform(0.2, 10.0, n)
    VR[0, :] = VL[0, :]                       # |du| < eps branch
    VL[1, ns:ns + dim], VR[1, ns:ns + dim] = 8.0, 8.5          # supersonic to the right / left: upwind overrides
    VL[2, ns:ns + dim], VR[2, ns:ns + dim] = -8.0, -8.5
    VL[3, ns:ns + dim], VR[3, ns:ns + dim] = 2.5, -2.5         # strong compression
    return VL, VR


def ref_riemann(lib, dim, ns, direction, VL, VR, th):
    neq = dim + 1 + ns
    F1, F2 = (C.c_double * neq)(), (C.c_double * neq)()
    rc = lib.ref_riemann_point_fc(dim, ns, direction, (C.c_double * neq)(*VL), (C.c_double * neq)(*VR),
                                  *[C.c_double(x) for x in th], F1, F2)
    assert rc == 0
    return np.array(F1[:]), np.array(F2[:])


def main():
    orc.build()
    lib = C.CDLL(REF_SO)
    lib.ref_riemann_point_fc.restype = C.c_int
    rng = np.random.default_rng(20261017)
    out = {}
    v7 = pp7_inputs(rng)
    o7 = np.zeros((len(v7), 15))
    for i, v in enumerate(v7):
        buf = (C.c_double * 15)()
        lib.ref_path_points7((C.c_double * 16)(*v), buf)
        o7[i] = buf[:]
    out["pp7_in"], out["pp7_out"] = v7, o7
    v8 = pp8_inputs(rng)
    o8 = np.zeros((len(v8), 16))
    for i, v in enumerate(v8):
        buf = (C.c_double * 16)()
        lib.ref_path_points8((C.c_double * 24)(*v), buf)
        o8[i] = buf[:]
    out["pp8_in"], out["pp8_out"] = v8, o8
    ns = 2
    for dim in (2, 3):
        for d in range(dim):
            VL, VR = riemann_inputs(rng, dim, ns)
            key = f"rp_fc{dim}d{d}"
            TH, F1s, F2s = [], [], []
            for n in range(VL.shape[0]):
                tl = orc.side_thermo(orc.FOUR_EQN_CONSERVATIVE, dim, ns, GAMMA_R, VL[n])
                tr = orc.side_thermo(orc.FOUR_EQN_CONSERVATIVE, dim, ns, GAMMA_R, VR[n])
                th = (tl[0], tr[0], tl[1], tr[1], tl[2], tr[2])
                F1, F2 = ref_riemann(lib, dim, ns, d, VL[n], VR[n], th)
                TH.append(th)
                F1s.append(F1)
                F2s.append(F2)
            out[key + "_VL"], out[key + "_VR"] = VL, VR
            out[key + "_thermo"], out[key + "_F_HLLC"], out[key + "_F_HYB"] = np.array(TH), np.array(F1s), np.array(F2s)
    out["gamma_R"] = np.array(GAMMA_R)
    path = os.path.join(HERE, "four_eqn_kernels.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("pp")})


if __name__ == "__main__":
    main()
