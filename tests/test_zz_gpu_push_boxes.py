"""hb2_push_boxes_dev on one GPU: the boxes of a periodic level live in separate state arrays on the same device and
store their six- (or four-) wide halo slabs straight into each other (what NavierStokesLevel does between GPUs over
CUDA IPC / NVLink); with the directions a box owns alone filled locally, every ghost cell -- faces, edges, corners --
must equal the periodic image bit for bit (the copy xfer::RefineSchedule::fillData makes,
RungeKuttaLevelIntegrator.cpp:1568/1701)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,N,world,ghosts", [(3, (24, 14, 16), 2, 6), (3, (16, 24, 12), 4, 6), (3, (12, 16, 24), 8, 6),
                                                (2, (24, 16), 2, 6), (3, (16, 16, 24), 8, 4)])
def test_push_boxes_fill_every_ghost(dim, N, world, ghosts):
    import torch

    from hamers_b200 import abi
    from hamers_b200.level import BoxDecomposition, oneshot_schedule, push_boxes_of

    g = ghosts
    rng = np.random.default_rng(3)
    neq = dim + 2
    full = rng.standard_normal((neq,) + tuple(reversed(N)))
    decs = [BoxDecomposition(dim, N, world, r) for r in range(world)]
    dx = tuple(1.0 / n for n in N)
    plan = abi.Plan(dim, decs[0].n, species_gamma=(1.4,), dx=dx, math=abi.MATH_EXACT, num_ghosts=g).use_torch_stream()
    try:
        states = []
        for dec in decs:
            U = torch.full((neq,) + tuple(x + 2 * g for x in reversed(dec.n)), float("nan"), dtype=torch.float64, device="cuda")
            box = tuple(slice(dec.lo[a], dec.lo[a] + dec.n[a]) for a in reversed(range(dim)))
            U[(slice(None),) + tuple(slice(g, -g) for _ in range(dim))] = torch.from_numpy(
                np.ascontiguousarray(full[(slice(None),) + box])).cuda()
            states.append(U)
        for dec, U in zip(decs, states):
            boxes, peers, shifts = push_boxes_of(dec, g)
            plan.push_boxes(U, plan.peer_box_table(boxes, [states[p].data_ptr() for p in peers], shifts))
        want = np.pad(full, [(0, 0)] + [(g, g)] * dim, mode="wrap")
        for dec, U in zip(decs, states):
            mask = oneshot_schedule(dec, neq, g)[2]
            if mask:
                plan.fill_ghosts_periodic(U, mask)
            box = tuple(slice(dec.lo[a], dec.lo[a] + dec.n[a] + 2 * g) for a in reversed(range(dim)))
            assert np.array_equal(U.cpu().numpy(), want[(slice(None),) + box])
    finally:
        plan.close()


def test_push_boxes_rejects_images_outside_the_ghost_box():
    import torch

    from hamers_b200 import abi

    plan = abi.Plan(3, (12, 12, 12), species_gamma=(1.4,), dx=(0.1, 0.1, 0.1), math=abi.MATH_EXACT, num_ghosts=6).use_torch_stream()
    try:
        U = torch.zeros((5, 24, 24, 24), dtype=torch.float64, device="cuda")
        with pytest.raises(Exception):
            plan.push_boxes(U, plan.peer_box_table([((6, 0, 0), (12, 12, 12))], [U.data_ptr()], [(24, 0, 0)]))
    finally:
        plan.close()
