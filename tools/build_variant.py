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
obj = os.path.join(hb.BUILD, f"hb2_sweeps_fast_{tag}.o")
cmd = [hb.NVCC] + hb.COMMON + ["-DHB2_MATH=1", "-fmad=true"] + flags + ["-c", os.path.join(hb.CSRC, "hb2_sweeps.cu"), "-o", obj]
subprocess.check_call(cmd)
objs = [os.path.join(hb.BUILD, u[0]) for u in hb.UNITS if u[0] != "hb2_sweeps_fast.o"] + [obj]
so = os.path.join(ROOT, "hamers_b200", f"libhamers_b200_{tag}.so")
subprocess.check_call([hb.NVCC] + hb.ARCH + ["-shared", "-o", so] + objs)
print(so)
