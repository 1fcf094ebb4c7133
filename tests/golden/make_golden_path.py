#!/usr/bin/env python
"""Generate tests/golden/path_vectors.npz: whole-path outputs of the ORACLE (side fluxes, source, one fused SSP-RK3
stage; spectral radii and stable dt) on the seeded branch-coverage states of tests/common.py, for the three
interpolators.  The oracle's point kernels are pinned to the reference's own compiled functions
(tests/golden/point_kernels.npz); these vectors freeze everything around them, so that a change of compiler, flags or
of the oracle itself shows up on any machine (tests/test_oracle_pinned.py), and the GPU exact build can be compared
with committed numbers as well as with the live oracle (tests/test_gpu_parity.py).

    python tests/golden/make_golden_path.py
"""
import dataclasses
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import make_case  # noqa: E402
from hamers_b200 import problems as pb  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = [("ss3d", 0), ("fe2d", 0), ("ss2d", 1), ("fe3d", 2)]   # (case of tests/common.py, scheme)
DT = 7.5e-4


def vectors(name, scheme):
    desc, U = make_case(name, "random")
    desc = dataclasses.replace(desc, scheme=scheme)
    Q = pb.pad_periodic(U)
    F, S = orc.compute_flux_and_source(desc, Q, DT)
    Un = orc.advance_stage(desc, [1.0], [1.0], [Q], [F], [S])
    sr, dt = orc.spectral_radii_and_dt(desc, Q, include_ghosts=False)
    return desc, Q, F, S, Un, sr, dt


def main():
    orc.build()
    out = {}
    for name, scheme in CASES:
        desc, Q, F, S, Un, sr, dt = vectors(name, scheme)
        key = f"{name}_s{scheme}"
        for a in range(desc.dim):
            out[f"{key}_F{a}"] = F[a]
        out[f"{key}_S"] = S
        sl = (slice(None),) + tuple(slice(4, -4) for _ in range(desc.dim))
        out[f"{key}_U"] = np.ascontiguousarray(Un[sl])
        out[f"{key}_sr"] = np.append(sr, dt)
    path = os.path.join(HERE, "path_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
