"""GPU counterpart of tests/test_host_emu.py::test_emulated_bounds_check_follows_the_reference_per_direction: the five-eqn
c^2 check of the first-order fallback follows the reference per direction (x: Gamma p/rho + sum Y_i Psi_i > 0; y, z:
Gamma p/rho > 0 -- FlowModelBasicUtilitiesFiveEqnAllaire.cpp:6654-6678 vs 6968-6992, 7281-7305).  Species gammas
(1.0005, 3) make the two forms disagree on the faces where an interpolated volume fraction undershoots zero."""
import dataclasses

import numpy as np
import pytest

from common import assert_fast_parity, make_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("math", [0, 1])
def test_bounds_check_follows_the_reference_per_direction(math, oracle_lib, product_lib):
    import torch
    from hamers_b200 import abi
    from hamers_b200 import problems as pb

    desc, U = make_case("fe3d", "random")
    desc = dataclasses.replace(desc, gamma=(1.0005, 3.0))
    Q = pb.pad_periodic(U)
    dt = 1.0e-3
    Fo, So = oracle_lib.compute_flux_and_source(desc, Q, dt)
    plan = abi.Plan(desc.dim, desc.n, flow_model=desc.model, species_gamma=desc.gamma, dx=desc.dx, weno_p=desc.weno_p,
                    math=math).use_torch_stream()
    Qd = torch.from_numpy(np.ascontiguousarray(Q)).cuda()
    Fd = [torch.zeros((desc.neq,) + desc.side_shape(a), dtype=torch.float64, device="cuda") for a in range(desc.dim)]
    Sd = torch.zeros((desc.neq,) + desc.cell_shape, dtype=torch.float64, device="cuda")
    plan.compute_flux_and_source(Qd, dt, Fd, Sd)
    torch.cuda.synchronize()
    for a in range(desc.dim):
        if math == 0:
            assert np.array_equal(Fd[a].cpu().numpy(), Fo[a]), f"dir {a}"
        else:
            assert_fast_parity(Fd[a].cpu().numpy(), Fo[a], f"dir {a}", 5.0e-3)
    if math == 0:
        assert np.array_equal(Sd.cpu().numpy(), So)
    plan.close()
