#!/usr/bin/env python
"""Generate tests/golden/three_species_kernels.npz from oracle/_ref/libhamers_ref.so: the reference's OWN HLLC / HLLC-HLL
point kernels of the five-eqn (FlowModelRiemannSolverFiveEqnAllaireHLLC{,-HLL}.cpp) and four-eqn
(FlowModelRiemannSolverFourEqnConservativeHLLC{,-HLL}.cpp) flow models called with d_num_species = 3 (the kernels take the
species count as an argument), fed with the oracle's side thermodynamics.  Needs /root/reference; run in the build container:

    python tests/golden/make_golden_three_species.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402
import make_golden_four_eqn as mg4  # noqa: E402
from oracle import oracle as orc  # noqa: E402

GAMMA_FE = (1.6, 1.4, 1.25)
GAMMA_R_FC = (1.6, 1.4, 1.25, 0.7, 1.3, 1.0)      # species gammas, then species gas constants
NS = 3


def main():
    orc.build()
    lib = C.CDLL(mg4.REF_SO)
    lib.ref_riemann_point.restype = C.c_int
    lib.ref_riemann_point_fc.restype = C.c_int
    rng = np.random.default_rng(20261018)
    out = {"gamma_fe": np.array(GAMMA_FE), "gamma_R_fc": np.array(GAMMA_R_FC)}
    for dim in (2, 3):
        for d in range(dim):
            VL, VR = mg.riemann_inputs(rng, 1, dim, NS)
            # the interpolated volume fractions of a side must leave room for the last one
            VL[:, NS + dim + 1:] *= 0.5
            VR[:, NS + dim + 1:] *= 0.5
            key = f"rp_fe{dim}d{d}"
            F1s, F2s, vms, ths = [], [], [], []
            for a, b in zip(VL, VR):
                F1, F2, vm, th = mg.ref_riemann(lib, 1, dim, NS, GAMMA_FE, d, a, b)
                F1s.append(F1), F2s.append(F2), vms.append(vm), ths.append(th)
            out[key + "_VL"], out[key + "_VR"], out[key + "_thermo"] = VL, VR, np.array(ths)
            out[key + "_F_HLLC"], out[key + "_F_HYB"], out[key + "_vel_mid"] = np.array(F1s), np.array(F2s), np.array(vms)
            VL, VR = mg4.riemann_inputs(rng, dim, NS)
            key = f"rp_fc{dim}d{d}"
            TH, F1s, F2s = [], [], []
            for n in range(VL.shape[0]):
                tl = orc.side_thermo(orc.FOUR_EQN_CONSERVATIVE, dim, NS, GAMMA_R_FC, VL[n])
                tr = orc.side_thermo(orc.FOUR_EQN_CONSERVATIVE, dim, NS, GAMMA_R_FC, VR[n])
                th = (tl[0], tr[0], tl[1], tr[1], tl[2], tr[2])
                F1, F2 = mg4.ref_riemann(lib, dim, NS, d, VL[n], VR[n], th)
                TH.append(th), F1s.append(F1), F2s.append(F2)
            out[key + "_VL"], out[key + "_VR"], out[key + "_thermo"] = VL, VR, np.array(TH)
            out[key + "_F_HLLC"], out[key + "_F_HYB"] = np.array(F1s), np.array(F2s)
    path = os.path.join(HERE, "three_species_kernels.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, sorted(out)[:6])


if __name__ == "__main__":
    main()
