#!/usr/bin/env python
"""Stall-reason breakdown (warp-sampling) of one kernel from `ncu --page source --csv`, by opcode class."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
seen, tot, byop = set(), {}, {}
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] in seen or not r[0].startswith("0x"):
        continue
    seen.add(r[0])
    src = r[col["Source"]].split()
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    cat = "FP64" if op in ("DFMA", "DMUL", "DADD", "DSETP") else op
    for h in hdr:
        if h.startswith("stall_") and not h.endswith("(Not Issued)"):
            v = int(r[col[h]])
            tot[h] = tot.get(h, 0) + v
            byop.setdefault(h, {})
            byop[h][cat] = byop[h].get(cat, 0) + v
S = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
    top = sorted(byop[k].items(), key=lambda kv: -kv[1])[:6]
    print("%-24s %6.2f%%  " % (k, 100 * v / S), " ".join("%s:%.1f" % (a, 100 * b / S) for a, b in top))
