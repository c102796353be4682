"""Gradient known-answer tests on the CPU: the package's differentiable torch prologue chained with the host build of
the hand-written adjoint (csrc/dps_adjoint.cuh), against the reference's own autograd gradients stored in
tests/golden/grad_*.npz (which oracle/gen_golden.py cross-checked by finite differences on the reference).
Tolerance: 1e-6 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from tests import helpers

CASES = [('grad_ibb_heun', 'ibb'), ('grad_ibb_euler', 'ibb'), ('grad_ibb_rk4', 'ibb'), ('grad_k2a_heun', 'k2a'),
         # leaves that enter Y_dyn (x_d''): the gradient also flows through dL/dZsub, the fault re-reduction and inv(Y_dyn)
         ('grad_ibb_z_heun', 'ibb'), ('grad_k2a_z_heun', 'k2a')]


def loss_of(result, gold, names):
    idx = [names.index(str(c)) for c in gold['loss_channels']]
    out = result[:, :, idx]
    target = torch.from_numpy(gold['target'])[None, :, None] * torch.ones_like(out)
    return ((out - target) ** 2).mean(dim=(1, 2))


@pytest.mark.parametrize('case,grid', CASES)
def test_adjoint_matches_reference_autograd(case, grid):
    sim, gold, leaves = helpers.sim_from_gold(case, grid, device='cpu', requires_grad=True)
    names = [str(s) for s in gold['names']]
    result = helpers.run_emulated_autograd(sim)
    loss = loss_of(result, gold, names)
    assert np.allclose(loss.detach().numpy(), gold['loss'], rtol=1e-9, atol=0)
    loss.sum().backward()
    for (gname, attr), leaf in leaves.items():
        ref = gold['grad/gen|%s|%s' % (gname, attr)]
        mine = leaf.grad.real.numpy().reshape(-1)
        rel = np.abs(mine - ref) / (np.abs(ref).max() + 1e-300)
        assert rel.max() < 1e-6, (gname, attr, mine, ref)
