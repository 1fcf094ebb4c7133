#!/usr/bin/env python
"""Condense `python -m hamers_b200.build --force -v` output: one line per kernel (registers, spills, stack).
   python -m hamers_b200.build --force -v 2>&1 | python tools/ptxas_summary.py [filter-regex]"""
import re
import subprocess
import sys

txt = sys.stdin.read()
flt = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
cur = None
unit = 0
seen_first = set()
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "Used" in line:
        regs = re.search(r"Used (\d+) registers", line).group(1)
        name = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"hb2::\(anonymous namespace\)::", "", name)
        name = re.sub(r"hb2::", "", name)
        tag = "B" if name in seen_first else "A"
        seen_first.add(name)
        rec = f"[{tag}] {name:60s} regs={regs} {extra}"
        if not flt or flt.search(rec):
            print(rec)
        cur = None
        continue
    if cur and "bytes stack frame" in line:
        m2 = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        extra = f"stack={m2.group(1)} spill_st={m2.group(2)} spill_ld={m2.group(3)}"
