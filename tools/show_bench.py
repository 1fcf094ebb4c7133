#!/usr/bin/env python
"""One-line view of a bench.py JSON line: python tools/show_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = {n: round(v["avg_ms"], 3) for n, v in d["roofline"]["kernels"].items()}
print(f"gpus {d['n_gpus']}  {d['value'] / 1e9:.3f} Gcell/s  {d['ms_per_step']:.3f} ms/step  frac {d['roofline']['stage']['fp64_frac']:.3f}  {k}")
