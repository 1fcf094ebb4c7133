#!/usr/bin/env python
"""Multi-GPU parity of the box-partitioned Navier-Stokes level (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/multi_gpu_ns_check.py [--size 40] [--steps 2]

Every rank advances its box with six-wide halos stored straight into the neighbours' state arrays (hb2_push_boxes_dev over
CUDA IPC / NVLink; HB2_NS_PUSH=0: the single-phase NCCL schedule, which also fills the freshly set state); rank 0 also
advances the WHOLE level as one box and the gathered boxes are compared with it: bit-identical in the exact build (a box
boundary must be invisible), <= 1e-12 in the fast build.  Exit code 0 on success."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=40)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from hamers_b200 import abi
    from hamers_b200.ns_level import NavierStokesLevel

    N = (args.size, args.size - 8, args.size + 8)
    ax = [(np.arange(n) + 0.5) / n for n in N]
    X = np.meshgrid(*reversed(ax), indexing="ij")[::-1]
    rng = np.random.default_rng(4)
    rho = 1.0 + 0.2 * np.sin(2 * np.pi * sum(X)) + 0.01 * rng.standard_normal(X[0].shape)
    vel = [0.4 * np.cos(2 * np.pi * X[a]) + 0.01 * rng.standard_normal(X[0].shape) for a in range(3)]
    p = 1.0 + 0.1 * np.cos(2 * np.pi * X[0])
    U = np.stack([rho] + [rho * v for v in vel] + [p / 0.4 + 0.5 * rho * sum(v * v for v in vel)])
    dt = 2.0e-4
    kw = dict(species_gamma=1.4, species_R=1.0, species_mu=0.05, species_mu_v=0.02, species_c_p=3.5, species_Pr=0.72)
    ok = True
    for math in (abi.MATH_EXACT, abi.MATH_FAST):
        lvl = NavierStokesLevel(3, N, math=math, **kw)
        d = lvl.decomp
        box = (slice(None),) + tuple(slice(d.lo[a], d.lo[a] + d.n[a]) for a in reversed(range(3)))
        lvl.interior().copy_(torch.from_numpy(np.ascontiguousarray(U[box])))
        for _ in range(args.steps):
            lvl.rk_step(dt)
        mine = lvl.S[lvl.cur][lvl._interior_slices()].contiguous()
        gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, gathered, dst=0)
        if rank == 0:
            full = np.empty_like(U)
            for r in range(world):
                dr = type(d)(3, N, world, r)
                b = (slice(None),) + tuple(slice(dr.lo[a], dr.lo[a] + dr.n[a]) for a in reversed(range(3)))
                full[b] = gathered[r].cpu().numpy()
            one_lvl = NavierStokesLevel(3, N, math=math, distributed=False, **kw)
            one_lvl.interior().copy_(torch.from_numpy(U))
            for _ in range(args.steps):
                one_lvl.rk_step(dt)
            torch.cuda.synchronize()
            one = one_lvl.S[one_lvl.cur][one_lvl._interior_slices()].cpu().numpy()
            one_lvl.close()
            if math == abi.MATH_EXACT:
                same = np.array_equal(full, one)
                print(f"[multi_gpu_ns_check] world {world} push={lvl.push} exact: bit-identical = {same}, max diff {np.abs(full - one).max():.3e}")
                ok &= same
            else:
                err = float((np.abs(full - one) / (np.abs(one) + np.abs(one).max(axis=(1, 2, 3), keepdims=True))).max())
                print(f"[multi_gpu_ns_check] world {world} push={lvl.push} fast: max relative difference {err:.3e}")
                ok &= err <= 1e-12
        lvl.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
