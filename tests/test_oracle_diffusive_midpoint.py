"""SURVEY.md row f4, midpoint family, CPU side: the oracle of DiffusiveFluxReconstructorMidpointSixthOrder
(oracle/oracle_diffusive.c: orc_compute_diffusive_flux_midpoint).

* pins: its twelve kernels (staggered derivative at the midpoints, node derivative, node-to-midpoint interpolation,
  five-midpoint reconstruction), the side-diffusivity statements and the side term table against the reference's own code
  compiled verbatim (oracle/build_ref.py: midpoint_kernels) -- committed outputs in
  tests/golden/diffusive_midpoint_kernels.npz (generator tests/golden/make_golden_diffusive_midpoint.py) and, when oracle/_ref
  is present, live;
* physics: the divergence of the midpoint-reconstructed flux converges at sixth order to the exact viscous term, agrees with
  the node reconstructor to truncation error, conserves, and vanishes on a uniform state."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from hamers_b200 import problems as pb
from oracle import oracle as orc
from test_oracle_diffusive import GAMMA, TR, exact_viscous_divergence, flux_divergence, smooth_state

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "diffusive_midpoint_kernels.npz"))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libhamers_ref.so")
G = 6
SHAPES = {2: (9, 7), 3: (6, 5, 4)}


def _windows(a, axis, width):
    """sliding windows of `width` along numpy axis `axis`, window index last"""
    return np.lib.stride_tricks.sliding_window_view(a, width, axis=axis)


@pytest.mark.parametrize("dim", [2, 3])
def test_midpoint_kernels_match_reference_golden(dim, oracle_lib):
    n = SHAPES[dim]
    u = GOLD[f"u{dim}"]
    for d in range(dim):
        ax = dim - 1 - d
        dx_inv = 1.0 / (0.1 + 0.07 * d)
        dmid, interp, dnode = GOLD[f"dmid{dim}d{d}"], GOLD[f"interp{dim}d{d}"], GOLD[f"dnode{dim}d{d}"]
        # midpoint m (array index m + G along d) from nodes m-3 .. m+2: windows of six starting at array index m + G - 3
        w6 = _windows(u, ax, 6)
        got_d = np.full(dmid.shape, np.nan)
        got_i = np.full(dmid.shape, np.nan)
        it = np.nditer(w6[..., 0], flags=["multi_index"])
        for _ in it:
            idx = it.multi_index
            win = w6[idx]
            dm, im, _ = oracle_lib.mid_point_kernels(win, dx_inv, [0.0] * 5, 0.0)
            tgt = list(idx)
            tgt[ax] += 3
            got_d[tuple(tgt)], got_i[tuple(tgt)] = dm, im
        assert np.array_equal(got_d, dmid, equal_nan=True), f"staggered derivative, dir {d}"
        assert np.array_equal(got_i, interp, equal_nan=True), f"interpolation, dir {d}"
        assert np.isfinite(dmid).sum() == dmid.size // dmid.shape[ax] * (dmid.shape[ax] - 6)     # all but three on each side
        # node derivative: the node oracle's seven-point kernel is the same formula (pinned there against the node class);
        # here against the midpoint class's own copy of it
        w7 = _windows(u, ax, 7)
        got_n = np.full(dnode.shape, np.nan)
        L = orc.lib()
        L.orc_diff_first_derivative.restype = C.c_double
        it = np.nditer(w7[..., 0], flags=["multi_index"])
        for _ in it:
            idx = it.multi_index
            tgt = list(idx)
            tgt[ax] += 3
            got_n[tuple(tgt)] = L.orc_diff_first_derivative((C.c_double * 7)(*w7[idx]), C.c_double(dx_inv))
        assert np.array_equal(got_n, dnode, equal_nan=True), f"node derivative, dir {d}"
        # reconstruction: face i (array index i) from the midpoint fluxes i-2 .. i+2 (array indices i + G - 2 ..)
        Fm, face = GOLD[f"Fm{dim}d{d}"], GOLD[f"face{dim}d{d}"]
        w5 = _windows(Fm, ax, 5)
        inner = [slice(G, -G)] * dim
        inner[ax] = slice(G - 2, G - 2 + n[d] + 1)
        sub = w5[tuple(inner)]
        got_f = np.zeros(face.shape)
        it = np.nditer(sub[..., 0], flags=["multi_index"])
        for _ in it:
            idx = it.multi_index
            got_f[idx] = 0.0 + oracle_lib.mid_point_kernels([0.0] * 6, 1.0, sub[idx], 3.0e-3)[2]
        assert np.array_equal(got_f, face), f"reconstruction, dir {d}"


@pytest.mark.parametrize("dim", [2, 3])
def test_side_diffusivities_and_term_table_match_reference_golden(dim, oracle_lib):
    vin = GOLD[f"side_in{dim}"]
    for d in range(dim):
        got = np.array([oracle_lib.mid_side_diffusivities(dim, d, v[0], v[1], v[2], v[3:3 + dim]) for v in vin])
        assert np.array_equal(got, GOLD[f"side_out{dim}d{d}"]), d
    tab = GOLD[f"side_terms{dim}"]
    nterms = 0
    for f in range(dim):
        for d in range(dim):
            for e in range(dim + 2):
                terms = oracle_lib.mid_side_terms(dim, f, d, e)
                want = [int(x) for x in tab[f, d, e] if x >= 0]
                assert [t[1] for t in terms] == want, (f, d, e)
                # the variables are those of the cell-data table (one function of the reference serves both families)
                n, var, _ = C.c_int(), (C.c_int * 4)(), (C.c_int * 4)()
                orc.lib().orc_diff_terms(dim, f, d, e, C.byref(n), var, (C.c_int * 4)())
                assert [t[0] for t in terms] == list(var[:n.value])
                nterms += len(terms)
    assert nterms == (45 if dim == 3 else 18)


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference)")
def test_midpoint_kernels_match_reference_live(oracle_lib):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_diffusive_midpoint as mg

    lib = C.CDLL(REF_SO)
    rng = np.random.default_rng(123)
    n = (5, 4, 3)
    u = np.ascontiguousarray(rng.standard_normal(mg.ghost_shape(n)) * 37.0)
    for d in range(3):
        ax = 2 - d
        ref_d, ref_i = mg.ref_kernel(lib, 0, 3, d, u, n, 7.5), mg.ref_kernel(lib, 2, 3, d, u, n, 7.5)
        w6 = _windows(u, ax, 6)
        it = np.nditer(w6[..., 0], flags=["multi_index"])
        for _ in it:
            idx = it.multi_index
            tgt = list(idx)
            tgt[ax] += 3
            dm, im, _ = oracle_lib.mid_point_kernels(w6[idx], 7.5, [0.0] * 5, 0.0)
            assert dm == ref_d[tuple(tgt)] and im == ref_i[tuple(tgt)]
    for dim in (2, 3):
        for d in range(dim):
            for _ in range(100):
                v = np.concatenate([rng.uniform(0.001, 1.0, 3), rng.uniform(-5, 5, 3)])
                buf = (C.c_double * 8)()
                lib.ref_mid_side_diffusivities(dim, d, (C.c_double * 6)(*v), buf)
                got = oracle_lib.mid_side_diffusivities(dim, d, v[0], v[1], v[2], v[3:3 + dim])
                assert np.array_equal(got, np.array(buf[:len(got)]))


def _run(dim, N, dt=1.0e-3):
    U, rho, vel, p = smooth_state(dim, N)
    desc = orc.PatchDesc(dim=dim, n=N, gamma=(GAMMA,), dx=tuple(1.0 / n for n in N))
    Q = pb.pad_periodic(U, orc.GD)
    F = orc.compute_diffusive_flux_midpoint(desc, TR, Q, dt)
    return desc, Q, F, flux_divergence(desc, F, dt), exact_viscous_divergence(dim, rho, vel, p)


@pytest.mark.parametrize("dim", [2, 3])
def test_midpoint_flux_divergence_converges_at_sixth_order(dim):
    errs = []
    for n in (16, 32):
        N = (n, n + 4, n + 2)[:dim]
        _, _, _, div, exact = _run(dim, N)
        assert np.abs(div[0]).max() == 0.0                         # no diffusive mass flux
        errs.append(np.abs(div[1:] - exact[1:]).max())
    order = np.log2(errs[0] / errs[1])
    assert errs[1] < 1.0e-4 * np.abs(exact[1:]).max() and order > 5.3, (errs, order)


@pytest.mark.parametrize("dim", [2, 3])
def test_midpoint_flux_structure(dim):
    N = (10, 12, 9)[:dim]
    desc, Q, F, div, _ = _run(dim, N)
    scale = np.abs(div).max()
    assert np.abs(div.reshape(desc.neq, -1).sum(axis=1)).max() < 1.0e-11 * scale * div[0].size      # telescoping sum
    for a in range(dim):
        ax = dim - a
        assert np.array_equal(np.take(F[a], 0, axis=ax), np.take(F[a], -1, axis=ax))               # periodic wrap
        assert np.array_equal(F[a][0], np.zeros_like(F[a][0])) and not np.signbit(F[a][0]).any()      # +0.0 mass flux
    # the two families discretise the same flux: they differ by truncation error only
    Fn = orc.compute_diffusive_flux(desc, TR, Q, 1.0e-3)
    for a in range(dim):
        assert np.abs(F[a] - Fn[a]).max() < 2.0e-2 * np.abs(Fn[a]).max()
        assert not np.array_equal(F[a], Fn[a])
    # only the cells within 5 of the interior are read
    Qp = Q.copy()
    outer = np.ones(Q.shape[1:], dtype=bool)
    outer[(slice(1, -1),) * dim] = False
    Qp[:, outer] = np.nan
    F2 = orc.compute_diffusive_flux_midpoint(desc, TR, Qp, 1.0e-3)
    assert all(np.array_equal(F2[a], F[a]) for a in range(dim))
    U0 = np.ones((desc.neq,) + desc.cell_shape) * np.array([1.3, 0.2, -0.4, 0.7, 5.0][:dim + 1] + [5.0]).reshape((-1,) + (1,) * dim)
    F0 = orc.compute_diffusive_flux_midpoint(desc, TR, pb.pad_periodic(U0, orc.GD), 1.0e-3)
    assert all(np.abs(f).max() == 0.0 for f in F0)
