#!/usr/bin/env python
"""Generate tests/golden/diffusive_midpoint_kernels.npz from oracle/_ref/libhamers_ref.so: the reference's OWN kernels of
DiffusiveFluxReconstructorMidpointSixthOrder (staggered derivative at the midpoints, derivative at the nodes, node-to-midpoint
interpolation, five-midpoint flux reconstruction; DiffusiveFluxReconstructorMidpointSixthOrder.cpp:68-1799), its
side-diffusivity statements and its side term table (FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:2581-2777, 2799-3657),
compiled verbatim by oracle/build_ref.py: midpoint_kernels.  Needs /root/reference; run in the build container:

    python tests/golden/make_golden_diffusive_midpoint.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libhamers_ref.so")
G = 6
SHAPES = {2: (9, 7), 3: (6, 5, 4)}          # interior cells (x, y[, z])


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def ghost_shape(n):
    return tuple(x + 2 * G for x in reversed(n))


def mid_shape(n, d):
    e = [x + 2 * G for x in n]
    e[d] += 1
    return tuple(reversed(e))


def ref_kernel(lib, kind, dim, d, u, n, dx_inv):
    """kind 0: derivative at midpoints, 1: derivative at nodes, 2: interpolation to midpoints (NaN where not computed)"""
    out = np.full(ghost_shape(n) if kind == 1 else mid_shape(n, d), np.nan)
    lib.ref_mid_kernel(kind, dim, d, G, dptr(u), (C.c_int * 3)(*(list(n) + [1] * (3 - dim))), C.c_double(dx_inv), dptr(out))
    return out


def ref_reconstruct(lib, dim, d, Fm, n, dt):
    e = list(n)
    e[d] += 1
    out = np.zeros(tuple(reversed(e)))
    lib.ref_mid_reconstruct(dim, d, G, dptr(Fm), (C.c_int * 3)(*(list(n) + [1] * (3 - dim))), C.c_double(dt), dptr(out))
    return out


def main():
    lib = C.CDLL(REF_SO)
    rng = np.random.default_rng(20261018)
    out = {}
    for dim, n in SHAPES.items():
        u = np.ascontiguousarray(rng.standard_normal(ghost_shape(n)) * 10.0 ** rng.integers(-2, 3))
        out[f"u{dim}"] = u
        for d in range(dim):
            dx_inv = 1.0 / (0.1 + 0.07 * d)
            for kind, tag in ((0, "dmid"), (1, "dnode"), (2, "interp")):
                out[f"{tag}{dim}d{d}"] = ref_kernel(lib, kind, dim, d, u, n, dx_inv)
            Fm = np.ascontiguousarray(rng.standard_normal(mid_shape(n, d)))
            out[f"Fm{dim}d{d}"] = Fm
            out[f"face{dim}d{d}"] = ref_reconstruct(lib, dim, d, Fm, n, 3.0e-3)
        vin = np.concatenate([rng.uniform(0.01, 0.2, (200, 3)), rng.uniform(-3, 3, (200, 3))], axis=1)
        out[f"side_in{dim}"] = vin
        for d in range(dim):
            res = np.zeros((200, 8 if dim == 3 else 7))
            for i, v in enumerate(vin):
                buf = (C.c_double * 8)()
                lib.ref_mid_side_diffusivities(dim, d, (C.c_double * 6)(*v), buf)
                res[i] = buf[:res.shape[1]]
            out[f"side_out{dim}d{d}"] = res
        tab = -np.ones((dim, dim, dim + 2, 4), dtype=np.int64)
        for f in range(dim):
            for d in range(dim):
                for e in range(dim + 2):
                    sd = (C.c_int * 4)()
                    m = lib.ref_diff_side_terms(dim, f, d, e, sd)
                    assert m >= 0
                    tab[f, d, e, :m] = sd[:m]
        out[f"side_terms{dim}"] = tab
    path = os.path.join(HERE, "diffusive_midpoint_kernels.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
