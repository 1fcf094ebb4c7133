import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hamers_b200 import abi, problems as pb
n = 512
sizes = [32] * 16
zlo = [sum(sizes[:k]) for k in range(len(sizes))]
boxes = [((0, 0, zlo[k]), (n, n, zlo[k] + sizes[k])) for k in range(len(sizes))]
lvl = abi.DeviceLevel(3, boxes, (n, n, n), flow_model=abi.SINGLE_SPECIES, species_gamma=(1.4,), dx=(2.0 / n,) * 3, math=abi.MATH_FAST)
U, _, gam = pb.convergence_single_species(3, 64)
U = np.tile(U, (1, 8, 8, 8))
arrs = []
for k in range(len(sizes)):
    t = torch.zeros((5, sizes[k] + 8, n + 8, n + 8), dtype=torch.float64).pin_memory()
    t[:, 4:-4, 4:-4, 4:-4].copy_(torch.from_numpy(U[:, zlo[k]:zlo[k] + sizes[k]]))
    arrs.append(t.numpy())
lvl.upload(arrs)
dt = 0.001 * 2.0 / n
def run(tag, bg):
    lvl.advance(dt); lvl.synchronize()
    side = torch.cuda.Stream()
    hbuf = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    dbuf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    dbuf2 = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    hbuf2 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    if bg:
        with torch.cuda.stream(side):
            for _ in range(40):
                if bg in ("h2d", "both"): dbuf.copy_(hbuf, non_blocking=True)
        if bg in ("d2h", "both"):
            side2 = torch.cuda.Stream()
            with torch.cuda.stream(side2):
                for _ in range(40): hbuf2.copy_(dbuf2, non_blocking=True)
    t0 = time.perf_counter()
    for _ in range(3): lvl.advance(dt)
    lvl.synchronize()
    el = (time.perf_counter() - t0) / 3
    torch.cuda.synchronize()
    print(tag, f"{1e3*el:.2f} ms/step", flush=True)
run("no background copies", None)
run("background H2D", "h2d")
run("background D2H", "d2h")
run("background both", "both")
run("no background copies", None)
lvl.close()
