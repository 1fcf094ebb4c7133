"""Generates tests/golden/diffusive_kernels.npz from oracle/_ref/libhamers_ref.so, i.e. from the reference's OWN
DiffusiveFluxReconstructorNodeSixthOrder kernels and statements compiled by oracle/build_ref.py (diffusive_kernels).
Run in the build container (needs /root/reference); the fixture is committed, this script documents how it was made.

  u{dim}d            (ghost box, 6 ghosts) random cell data, n = (5, 4, 3)[:dim]
  der{dim}d{dir}     computeFirstDerivativesIn{X,Y,Z}(u) over the base class's range (NaN where not written)
  rec{dim}d{dir}     reconstructFlux{X,Y,Z}(u as node flux) into a zero-filled side array, dt = 0.37
  terms{dim}d        (flux dir, derivative dir, equation, term, 2) = (variable, diffusivity index) in the reference's order of
                     accumulation, -1 padded: FlowModelDiffusiveFluxUtilitiesSingleSpecies::getCellDataOfDiffusiveFluxVariables
                     ForDerivative / getCellDataOfDiffusiveFluxDiffusivities compiled verbatim (variable dim = temperature)
  dt{dim}d_in (n, 9) = mu, mu_v, kappa, c_p, rho, dx[3], max acoustic sum;  dt{dim}d_out (n, 3) = MAX_DIFFUSIVITY, diffusive
                     spectral radius, stable dt (FlowModelSingleSpecies.cpp:4661-4665; NavierStokes.cpp:884-893, 1083-1091)
  point{dim}d_in (n, 11) = gamma, c_v, rho, p, c_p, mu, Pr, mu_v, u, v, w;  point{dim}d_out (n, 2 + 13 | 10) = T, kappa, D_xx
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
N = (5, 4, 3)
DX_INV = 7.3
DT = 0.37


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhamers_ref.so"))
    P = C.POINTER(C.c_double)
    rng = np.random.default_rng(2026)
    out = {}
    for dim in (2, 3):
        n = N[:dim]
        shape = tuple(x + 12 for x in reversed(n))
        u = rng.standard_normal(shape)
        out[f"u{dim}d"] = u
        nn = (C.c_int * 3)(*(list(n) + [1] * (3 - dim)))
        for d in range(dim):
            der = np.full(shape, np.nan)
            lib.ref_diff_derivative(dim, d, u.ctypes.data_as(P), nn, C.c_double(DX_INV), der.ctypes.data_as(P))
            out[f"der{dim}d{d}"] = der
            fs = list(n)
            fs[d] += 1
            rec = np.zeros(tuple(reversed(fs)))
            lib.ref_diff_reconstruct(dim, d, u.ctypes.data_as(P), nn, C.c_double(DT), rec.ctypes.data_as(P))
            out[f"rec{dim}d{d}"] = rec
        # the term tables, from the reference's own getCellDataOfDiffusiveFluxVariablesForDerivative / ...Diffusivities
        terms = np.full((dim, dim, dim + 2, 4, 2), -1, dtype=np.int64)
        for f in range(dim):
            for d in range(dim):
                for e in range(dim + 2):
                    v, k = (C.c_int * 4)(), (C.c_int * 4)()
                    cnt = lib.ref_diff_terms(dim, f, d, e, v, k)
                    assert 0 <= cnt <= 4
                    for i in range(cnt):
                        terms[f, d, e, i] = (v[i], k[i])
        out[f"terms{dim}d"] = terms
        pin = np.abs(rng.standard_normal((200, 11))) * 10.0 ** rng.uniform(-2, 2, (200, 11)) + 1.0e-3
        pin[:, 0] = rng.uniform(1.1, 1.7, 200)
        pin[:, 8:11] *= rng.choice([-1.0, 1.0], (200, 3))
        pout = []
        for v in pin:
            o = (C.c_double * 15)()
            lib.ref_diff_point(dim, (C.c_double * 11)(*v), o)
            pout.append(list(o)[:2 + (13 if dim == 3 else 10)])
        out[f"point{dim}d_in"], out[f"point{dim}d_out"] = pin, np.array(pout)
        # MAX_DIFFUSIVITY, diffusive spectral radius and stable dt (oracle/build_ref.py: diffusive_dt_statements)
        din = np.abs(rng.standard_normal((200, 9))) * 10.0 ** rng.uniform(-3, 2, (200, 9)) + 1.0e-6
        dout = []
        for v in din:
            o = (C.c_double * 3)()
            lib.ref_diff_dt_point(dim, (C.c_double * 9)(*v), o)
            dout.append(list(o))
        out[f"dt{dim}d_in"], out[f"dt{dim}d_out"] = din, np.array(dout)
    np.savez_compressed(os.path.join(HERE, "diffusive_kernels.npz"), **out)
    print("wrote diffusive_kernels.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
