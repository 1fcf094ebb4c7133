#!/usr/bin/env python
"""Dynamic opcode mix + stall samples of one kernel from `ncu --page source --csv` output.
   ncu -i rep --page source --csv --kernel-name regex:X --launch-count 1 > src.csv; python tools/ncu_opmix.py src.csv"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
ops = defaultdict(lambda: [0, 0, 0])
tot = 0
samples = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    src = r[col["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0].rstrip(";")
    n = int(r[col["Instructions Executed"]])
    s = int(r[col["Warp Stall Sampling (All Samples)"]])
    ops[op][0] += n
    ops[op][1] += s
    ops[op][2] += 1
    tot += n
    samples += s
    lines.append((s, n, src))
print(f"total warp instructions {tot}, stall samples {samples}, static {len(lines)}")
for op, (n, s, c) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"  {op:10s} {n:12d} {100.0 * n / tot:6.2f}%  samples {100.0 * s / max(samples, 1):6.2f}%  static {c}")
fp = sum(v[0] for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"FP64-pipe share of issued instructions: {100.0 * fp / tot:.1f}%")
if len(sys.argv) > 2:
    print("top stall lines:")
    for s, n, src in sorted(lines, reverse=True)[:int(sys.argv[2])]:
        print(f"  {s:6d} {n:10d}  {src}")
