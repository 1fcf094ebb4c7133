#!/usr/bin/env python
"""Build a tuning variant of the library: only the fast sweeps are recompiled with extra flags and linked with the
other objects of the last full build into hamers_b200/libhamers_b200_<tag>.so; select it at run time with
HAMERS_B200_LIB=<path> (tools only; the product loads libhamers_b200.so).
    python tools/build_variant.py <tag> [nvcc flags ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hamers_b200 import build as hb  # noqa: E402

tag, flags = sys.argv[1], sys.argv[2:]
# HB2_VARIANT_UNIT=hb2_sweeps_fast_ld.o (or ..._fast_z.o) rebuilds that unit instead of the WCNS5-JS one
unit = os.environ.get("HB2_VARIANT_UNIT", "hb2_sweeps_fast.o")
base = [u for u in hb.UNITS if u[0] == unit][0]
obj = os.path.join(hb.BUILD, unit.replace(".o", f"_{tag}.o"))
cmd = [hb.NVCC] + hb.COMMON + base[2] + flags + ["-c", os.path.join(hb.CSRC, base[1]), "-o", obj]
subprocess.check_call(cmd)
objs = [os.path.join(hb.BUILD, u[0]) for u in hb.UNITS if u[0] != unit] + [obj]
so = os.path.join(ROOT, "hamers_b200", f"libhamers_b200_{tag}.so")
subprocess.check_call([hb.NVCC] + hb.ARCH + ["-shared", "-o", so] + objs)
print(so)
