"""Build libhamers_b200.so (the C-ABI product library) in-tree with nvcc for sm_100a.

    python -m hamers_b200.build [--force]

Eight translation units: the sweeps are compiled six times (exact: -fmad=false, reference operation
order, once per nonlinear interpolator WCNS5-JS / WCNS5-Z / WCNS6-LD; fast: FMA contraction + reciprocal
sharing, one per interpolator) and linked with the ABI layer and the
diffusive-flux unit (-fmad=false).  nvcc
cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(ROOT, "build")
SO = os.path.join(HERE, "libhamers_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include")]

UNITS = [
    ("hb2_sweeps_exact.o", "hb2_sweeps.cu", ["-DHB2_MATH=0", "-fmad=false"]),
    ("hb2_sweeps_exact_z.o", "hb2_sweeps.cu", ["-DHB2_MATH=0", "-DHB2_SCHEME=1", "-fmad=false"]),
    ("hb2_sweeps_exact_ld.o", "hb2_sweeps.cu", ["-DHB2_MATH=0", "-DHB2_SCHEME=2", "-fmad=false"]),
    ("hb2_sweeps_fast.o", "hb2_sweeps.cu", ["-DHB2_MATH=1", "-fmad=true"] + os.environ.get("HB2_FAST_FLAGS", "").split()),
    ("hb2_sweeps_fast_z.o", "hb2_sweeps.cu", ["-DHB2_MATH=1", "-DHB2_SCHEME=1", "-fmad=true"]),
    ("hb2_sweeps_fast_ld.o", "hb2_sweeps.cu", ["-DHB2_MATH=1", "-DHB2_SCHEME=2", "-fmad=true"]),
    ("hb2_abi.o", "hb2_abi.cu", ["-fmad=false"]),
    ("hb2_diffusive.o", "hb2_diffusive.cu", ["-fmad=false"]),
    ("hb2_amr.o", "hb2_amr.cu", ["-fmad=false"]),
    ("hb2_level.o", "hb2_level.cu", ["-fmad=false"]),
]
DEPS = ["hb2_diffusive_march.cuh", "hb2_core.cuh", "hb2_fast.cuh", "hb2_sweep.cuh", "hb2_sensor.cuh", "hb2_ops.h", "hb2_sweeps.cu", "hb2_abi.cu", "hb2_diffusive.cu", "hb2_diffusive.cuh", "hb2_amr.cu", "hb2_amr.cuh", "hb2_level.cu", os.path.join(ROOT, "include", "hamers_b200.h")]


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    newest = max(_mtime(d if os.path.isabs(d) else os.path.join(CSRC, d)) for d in DEPS)
    return newest > _mtime(SO)


def unit_deps(src: str) -> set:
    """The files a translation unit depends on: its source and, recursively, every quoted include found under csrc/ or
    include/ (so that a change of one header rebuilds only the units that see it)."""
    import re

    seen, todo = set(), [os.path.join(CSRC, src)]
    while todo:
        f = todo.pop()
        if f in seen or not os.path.exists(f):
            continue
        seen.add(f)
        for inc in re.findall(r'#\s*include\s+"([^"]+)"', open(f).read()):
            for d in (os.path.dirname(f), CSRC, os.path.join(ROOT, "include")):
                if os.path.exists(os.path.join(d, inc)):
                    todo.append(os.path.join(d, inc))
                    break
    return seen


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stale = [u for u in UNITS if force or max(_mtime(d) for d in unit_deps(u[1])) > _mtime(os.path.join(BUILD, u[0]))]
    if not stale and os.path.exists(SO) and all(_mtime(os.path.join(BUILD, u[0])) <= _mtime(SO) for u in UNITS):
        return SO

    def compile_one(unit):
        obj, src, flags = unit
        cmd = [NVCC] + COMMON + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", os.path.join(BUILD, obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src} {flags}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        logs = list(ex.map(compile_one, stale))
    if verbose:
        for lg in logs:
            print(lg)
    link = [NVCC] + ARCH + ["-shared", "-o", SO] + [os.path.join(BUILD, u[0]) for u in UNITS]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return SO


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
