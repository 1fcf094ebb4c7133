"""hb2_level_* through the ctypes binding (abi.DeviceLevel): the device-resident multi-patch level and its pipelined
host-memory advance (uploads of later patches overlap the first stage of the earlier ones, downloads overlap the last stage)
against the one-patch entry point and the oracle level: bit-identical in the reference-order build."""
import numpy as np
import pytest

from hamers_b200 import problems as pb

pytestmark = pytest.mark.gpu
G = 4


def _patches(U, boxes, dim):
    """ghost-box host arrays of the patches, interior from the level array U, ghosts poisoned"""
    out = []
    for lo, hi in boxes:
        shape = (U.shape[0],) + tuple(hi[a] - lo[a] + 2 * G for a in reversed(range(dim)))
        a = np.full(shape, np.nan)
        a[(slice(None),) + (slice(G, -G),) * dim] = U[(slice(None),) + tuple(slice(lo[d], hi[d]) for d in reversed(range(dim)))]
        out.append(a)
    return out


@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("model,dim,N,cuts", [(0, 3, (24, 20, 32), [(0, 8), (8, 16), (16, 24), (24, 32)]),
                                            (1, 3, (16, 12, 20), [(0, 7), (7, 20)]), (0, 2, (40, 36), [(0, 9), (9, 20), (20, 36)])])
def test_pipelined_host_advance_matches_the_oracle_level(model, dim, N, cuts, math, oracle_lib, product_lib):
    from common import assert_fast_parity
    from hamers_b200 import abi

    U, dx, gam = pb.random_state(dim, N, model=model, seed=21, shock=False)
    mean = U.mean(axis=tuple(range(1, dim + 1)), keepdims=True)
    U = np.ascontiguousarray(mean + 0.2 * (U - mean))
    # slabs along the slowest direction, like a host box cut for the transfer pipeline
    boxes = [((0,) * (dim - 1) + (a,), tuple(N[:dim - 1]) + (b,)) for a, b in cuts]
    lvl = abi.DeviceLevel(dim, boxes, N, flow_model=model, species_gamma=gam, dx=dx, math=math)
    host = _patches(U, boxes, dim)
    dt = 2.0e-3 * min(dx)
    for _ in range(2):
        lvl.advance_host(host, dt)
    got = np.empty_like(U)
    for (lo, hi), a in zip(boxes, host):
        got[(slice(None),) + tuple(slice(lo[d], hi[d]) for d in reversed(range(dim)))] = a[(slice(None),) + (slice(G, -G),) * dim]
    desc = oracle_lib.PatchDesc(dim=dim, n=N, model=model, ns=len(gam), gamma=gam, dx=dx)
    want = U.copy()
    oracle_lib.level_advance(desc, N, want, dt, 2, nthreads=0)
    if math == 0:
        assert np.array_equal(got, want)
    else:
        assert_fast_parity(got, want, "two pipelined SSP-RK3 steps")
    # the device-resident route gives the same bits as the host route
    lvl2 = abi.DeviceLevel(dim, boxes, N, flow_model=model, species_gamma=gam, dx=dx, math=math)
    host2 = _patches(U, boxes, dim)
    lvl2.upload(host2)
    lvl2.advance(dt)
    lvl2.advance(dt)
    lvl2.download(host2)
    for a, b in zip(host, host2):
        inner = (slice(None),) + (slice(G, -G),) * dim
        assert np.array_equal(a[inner], b[inner])
    assert lvl.launch_count > 0
    lvl.close()
    lvl2.close()
