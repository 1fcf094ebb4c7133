#!/usr/bin/env python
"""Multi-GPU parity of the box-partitioned level (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/multi_gpu_check.py [--size 48] [--steps 2] [--model ss|fe]

Every rank advances its box of a periodic uniform level with the halo exchange of hamers_b200.level (NCCL P2P over
NVLink); rank 0 also advances the WHOLE level on its own GPU (single box, periodic ghost fill kernel) and the gathered
boxes are compared with it: bit-identical in the exact build (a box boundary must be invisible), <= 1e-12 in the fast
build.  Exit code 0 on success."""
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
    ap.add_argument("--size", type=int, default=48)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--model", default="ss")
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from hamers_b200 import abi
    from hamers_b200 import problems as pb
    from hamers_b200.level import UniformLevel

    N = (args.size, args.size - 8, args.size + 8)
    model = abi.SINGLE_SPECIES if args.model == "ss" else abi.FIVE_EQN_ALLAIRE
    U, dx, gam = pb.random_state(3, N, model=model, seed=4, shock=True)
    dt = 2.0e-4
    ok = True
    for math in (abi.MATH_EXACT, abi.MATH_FAST):
        lvl = UniformLevel(3, N, flow_model=model, species_gamma=gam, math=math)
        d = lvl.decomp
        box = (slice(None),) + tuple(slice(d.lo[a], d.lo[a] + d.n[a]) for a in reversed(range(3)))
        lvl.set_interior(U[box])
        dt_level = lvl.stable_dt(0.5)      # spectral radii per box + MAX all-reduce over the ranks (row f1)
        lvl.advance(dt, args.steps)
        mine = lvl.interior().contiguous()
        gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        # boxes have equal shapes (regular decomposition)
        dist.gather(mine, gathered, dst=0)
        if rank == 0:
            full = np.empty_like(U)
            for r in range(world):
                dr = type(d)(3, N, world, r)
                b = (slice(None),) + tuple(slice(dr.lo[a], dr.lo[a] + dr.n[a]) for a in reversed(range(3)))
                full[b] = gathered[r].cpu().numpy()
            # reference: the whole level as ONE box on this GPU (no torch.distributed inside)
            plan = abi.Plan(3, N, flow_model=model, species_gamma=gam, dx=lvl.dx, math=math).use_torch_stream()
            S = torch.from_numpy(pb.pad_periodic(U)).cuda()
            sr = torch.zeros(4, dtype=torch.float64, device="cuda")
            plan.max_wave_speed(S, sr)
            dt_one = 0.5 / float(sr[3])
            print(f"[multi_gpu_check] world {world} stable dt: level {dt_level:.17g}, single box {dt_one:.17g}")
            ok &= dt_level == dt_one
            for _ in range(args.steps):
                plan.advance_level(S, dt)
            torch.cuda.synchronize()
            one = S[:, 4:-4, 4:-4, 4:-4].cpu().numpy()
            plan.close()
            if math == abi.MATH_EXACT:
                same = np.array_equal(full, one)
                print(f"[multi_gpu_check] world {world} exact: bit-identical = {same}, max diff {np.abs(full - one).max():.3e}")
                ok &= same
            else:
                err = float((np.abs(full - one) / (np.abs(one) + np.abs(one).max(axis=(1, 2, 3), keepdims=True))).max())
                print(f"[multi_gpu_check] world {world} fast: max relative difference {err:.3e}")
                ok &= err <= 1e-12
        lvl.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
