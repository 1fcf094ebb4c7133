#!/usr/bin/env python
"""PCIe micro-benchmark: H2D alone, D2H alone, both at once on two streams (pinned memory).  python tools/pcie_duplex.py [GiB]"""
import sys
import time

import torch

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gib * 2 ** 30 / 8)
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, chunks=16):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = n // chunks
    for k in range(chunks):
        if h2d:
            with torch.cuda.stream(s1):
                d_in[k * m:(k + 1) * m].copy_(h_in[k * m:(k + 1) * m], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[k * m:(k + 1) * m].copy_(d_out[k * m:(k + 1) * m], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for _ in range(2):
    a, b, c = run(True, False), run(False, True), run(True, True)
print(f"{gib} GiB each way: H2D {gib * 1.0737 / a:.1f} GB/s, D2H {gib * 1.0737 / b:.1f} GB/s, both at once {c * 1e3:.1f} ms "
      f"(alone {a * 1e3:.1f} + {b * 1e3:.1f} ms): {gib * 1.0737 / c:.1f} GB/s per direction")
