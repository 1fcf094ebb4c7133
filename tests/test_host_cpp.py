"""The host-side C++ mirror of the reference interface (hamers_b200/host): same class / method names as
ConvectiveFluxReconstructor{,WCNS5_JS_HLLC_HLL} and the SAMRAI patch-data layout, marshalling into the C ABI.
The CPU part checks that it compiles against the SAMRAI shim with the reference's C++ dialect (-std=c++11) and links
against the product library; the GPU part runs it like HAMeRS would and compares with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from common import CASES, assert_fast_parity, interior, make_case
from hamers_b200 import problems as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "hamers_b200", "host")
EXE = os.path.join(ROOT, "tests", "host_cpp", "test_reconstructor")


def build_driver():
    srcs = [os.path.join(ROOT, "tests", "host_cpp", "test_reconstructor.cpp"),
            os.path.join(HOST, "ConvectiveFluxReconstructorB200.cpp")]
    deps = srcs + [os.path.join(HOST, "ConvectiveFluxReconstructorB200.hpp"), os.path.join(HOST, "samrai_shim.hpp"),
                   os.path.join(ROOT, "include", "hamers_b200.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(d) <= os.path.getmtime(EXE) for d in deps):
        return EXE
    libdir = os.path.join(ROOT, "hamers_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-Wextra", "-o", EXE] + srcs +
                          ["-L", libdir, "-lhamers_b200", "-Wl,-rpath," + libdir])
    return EXE


def test_host_classes_compile_and_link(product_lib):
    assert os.path.exists(build_driver())


# ---- DiffusiveFluxReconstructorNodeSixthOrder_B200 ----------------------------------------------------
DEXE = os.path.join(ROOT, "tests", "host_cpp", "test_diffusive")
TRANSPORT = dict(R=1.0, mu=0.05, mu_v=0.02, c_p=3.5, Pr=0.72)


def build_diffusive_driver():
    srcs = [os.path.join(ROOT, "tests", "host_cpp", "test_diffusive.cpp"), os.path.join(HOST, "DiffusiveFluxReconstructorB200.cpp"),
            os.path.join(HOST, "ConvectiveFluxReconstructorB200.cpp")]
    deps = srcs + [os.path.join(HOST, "DiffusiveFluxReconstructorB200.hpp"), os.path.join(HOST, "ConvectiveFluxReconstructorB200.hpp"),
                   os.path.join(HOST, "samrai_shim.hpp"), os.path.join(ROOT, "include", "hamers_b200.h")]
    if os.path.exists(DEXE) and all(os.path.getmtime(d) <= os.path.getmtime(DEXE) for d in deps):
        return DEXE
    libdir = os.path.join(ROOT, "hamers_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-Wextra", "-Werror", "-o", DEXE] + srcs +
                          ["-L", libdir, "-lhamers_b200", "-Wl,-rpath," + libdir])
    return DEXE


def run_diffusive_driver(tmp_path, dim, N, U6, dx, gamma, dt):
    fin, fout = str(tmp_path / "din.bin"), str(tmp_path / "dout.bin")
    with open(fin, "wb") as fh:
        fh.write(struct.pack("4i", dim, *(list(N) + [1] * (3 - dim))))
        t = TRANSPORT
        fh.write(struct.pack("10d", gamma, t["R"], t["mu"], t["mu_v"], t["c_p"], t["Pr"], *(list(dx) + [1.0] * (3 - dim)), dt))
        fh.write(np.ascontiguousarray(U6).tobytes())
    return subprocess.run([build_diffusive_driver(), fin, fout], capture_output=True, text=True), fout


def test_diffusive_host_class_compiles_links_and_refuses_to_run_without_a_device(product_lib, tmp_path):
    """-std=c++11 -Werror against the SAMRAI shim; without a CUDA device the class surfaces the library's message through
    TBOX_ERROR (no CPU fallback)."""
    import ctypes

    n = ctypes.c_int32()
    product_lib.hb2_device_count(ctypes.byref(n))
    U, dx, gam = pb.random_state(3, (6, 5, 4), seed=1, shock=False)
    r, _ = run_diffusive_driver(tmp_path, 3, (6, 5, 4), pb.pad_periodic(U, 6), dx, gam[0], 1.0e-3)
    assert "Print DiffusiveFluxReconstructorNodeSixthOrder_B200 object" in r.stdout
    if n.value == 0:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr, r.stdout + r.stderr
    else:
        assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("math", [0, 1, 10, 20])
@pytest.mark.parametrize("name", list(CASES))
def test_reconstructor_class_matches_oracle(name, math, oracle_lib, tmp_path):
    """math: 0 exact / 1 fast WCNS5_JS_HLLC_HLL; 10 = WCNS5_Z_HLLC_HLL, 20 = WCNS6_LD_HLLC_HLL (reference-order kernels)."""
    import dataclasses

    exe = build_driver()
    desc, U = make_case(name, "random")
    desc = dataclasses.replace(desc, scheme=math // 10)
    Q = pb.pad_periodic(U)
    dt = 7.5e-4
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as fh:
        n = list(desc.n) + [1] * (3 - desc.dim)
        fh.write(struct.pack("7i", desc.dim, n[0], n[1], n[2], desc.model, desc.ns, math))
        g = list(desc.gamma) + [0.0] * (4 - len(desc.gamma))
        dx = list(desc.dx) + [1.0] * (3 - desc.dim)
        fh.write(struct.pack("8d", *(g + dx + [dt])))
        fh.write(np.ascontiguousarray(Q).tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "d_constant_p = 2" in r.stdout
    if math == 20:
        assert "d_constant_alpha_tau = 35" in r.stdout
    math = math % 10
    out = np.fromfile(fout)
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt)
    Uo = oracle_lib.advance_stage(desc, [1.0], [1.0], [Q], [Fo], [So])
    pos = 0
    for a in range(desc.dim):
        k = Fo[a].size
        Fg = out[pos:pos + k].reshape(Fo[a].shape)
        pos += k
        if math == 0:
            assert np.array_equal(Fg, Fo[a]), f"dir {a}"
        else:
            assert_fast_parity(Fg, Fo[a], f"dir {a}")
    Sg = out[pos:pos + So.size].reshape(So.shape)
    pos += So.size
    Ug = out[pos:pos + Uo.size].reshape(Uo.shape)
    if math == 0:
        assert np.array_equal(Sg, So)
        assert np.array_equal(interior(desc, Ug), interior(desc, Uo))
    else:
        assert_fast_parity(Sg, So, "source")
        assert_fast_parity(interior(desc, Ug), interior(desc, Uo), "fused stage")


# ---- device-resident seam 2: RungeKuttaPatchStrategyB200 over a multi-patch level ---------------------------------------
PEXE = os.path.join(ROOT, "tests", "host_cpp", "test_patch_strategy")


def build_patch_strategy_driver():
    srcs = [os.path.join(ROOT, "tests", "host_cpp", "test_patch_strategy.cpp"), os.path.join(HOST, "RungeKuttaPatchStrategyB200.cpp"),
            os.path.join(HOST, "ConvectiveFluxReconstructorB200.cpp")]
    deps = srcs + [os.path.join(HOST, "RungeKuttaPatchStrategyB200.hpp"), os.path.join(HOST, "ConvectiveFluxReconstructorB200.hpp"),
                   os.path.join(HOST, "samrai_shim.hpp"), os.path.join(ROOT, "include", "hamers_b200.h")]
    if os.path.exists(PEXE) and all(os.path.getmtime(d) <= os.path.getmtime(PEXE) for d in deps):
        return PEXE
    libdir = os.path.join(ROOT, "hamers_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-Wextra", "-Werror", "-o", PEXE] + srcs +
                          ["-L", libdir, "-lhamers_b200", "-Wl,-rpath," + libdir])
    return PEXE


def run_patch_strategy_driver(tmp_path, dim, N, model, ns, math, nsteps, cuts, gam, R, dx, dt, U):
    fin, fout = str(tmp_path / "pin.bin"), str(tmp_path / "pout.bin")
    with open(fin, "wb") as fh:
        fh.write(struct.pack("11i", dim, *(list(N) + [1] * (3 - dim)), model, ns, math, nsteps, *(list(cuts) + [0] * (3 - dim))))
        fh.write(struct.pack("12d", *(list(gam) + [1.4] * (4 - len(gam))), *(list(R) + [1.0] * (4 - len(R))),
                             *(list(dx) + [1.0] * (3 - dim)), dt))
        fh.write(np.ascontiguousarray(U).tobytes())
    return subprocess.run([build_patch_strategy_driver(), fin, fout], capture_output=True, text=True), fout


def test_patch_strategy_class_compiles_and_links(product_lib):
    assert os.path.exists(build_patch_strategy_driver())


@pytest.mark.gpu
@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("model,dim,N,cuts", [(0, 3, (16, 16, 16), (8, 8, 8)), (0, 3, (20, 14, 12), (7, 9, 5)), (1, 2, (24, 20), (10, 13)),
                                            (2, 2, (18, 16), (5, 8))])
def test_device_resident_multi_patch_level_matches_the_oracle_level(model, dim, N, cuts, math, oracle_lib, product_lib, tmp_path):
    """RungeKuttaPatchStrategyB200 driven like RungeKuttaLevelIntegrator::advanceLevel over 2 x 2 (x 2) patches of unequal
    sizes (registered once, ghost fill and stages on the device copies, download at the end) against orc_level_advance on the
    same periodic level: bit-identical in the reference-order build, <= 1e-12 in the fast build; and the level's spectral
    radii / stable dt."""
    R = ()
    if model == 2:
        U, dx, gam, R = pb.random_state_four_eqn(dim, N, seed=9, shock=False)
    else:
        U, dx, gam = pb.random_state(dim, N, model=model, seed=9, shock=False)
    # tame the white noise: a few steps must stay well inside the physical range
    mean = U.mean(axis=tuple(range(1, dim + 1)), keepdims=True)
    U = np.ascontiguousarray(mean + 0.2 * (U - mean))
    ns = len(gam)
    dt, nsteps = 2.0e-3 * min(dx), 2
    r, fout = run_patch_strategy_driver(tmp_path, dim, N, model, ns, math, nsteps, cuts, gam, R, dx, dt, U)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "test_patch_strategy OK" in r.stdout
    desc = oracle_lib.PatchDesc(dim=dim, n=N, model=model, ns=ns, gamma=gam, R=R, dx=dx)
    ncomp = desc.ncomp
    raw = np.fromfile(fout, dtype=np.float64)
    got = raw[:ncomp * int(np.prod(N))].reshape((ncomp,) + tuple(reversed(N)))
    sr = raw[ncomp * int(np.prod(N)):]
    want = U.copy()
    oracle_lib.level_advance(desc, N, want, dt, nsteps, nthreads=0)
    if math == 0 or model == 2:
        assert np.array_equal(got, want)
    else:
        assert_fast_parity(got, want, "two SSP-RK3 steps of the multi-patch level")
    radii, dt_o = oracle_lib.spectral_radii_and_dt(desc, pb.pad_periodic(want), include_ghosts=False)
    if math == 0 or model == 2:
        assert np.array_equal(sr[:dim], radii)
    assert abs(sr[dim] - dt_o) <= 1e-12 * dt_o
