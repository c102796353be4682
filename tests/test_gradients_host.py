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
         ('grad_ibb_z_heun', 'ibb'), ('grad_k2a_z_heun', 'k2a'),
         # BASELINE config 4 as the reference runs it: H / D / X_d' leaves from the 0.5x..2x identification range, the full
         # 10 s, loss = per-sim MSE against the reference's own golden trajectory (ibb_model.npy)
         ('grad_ibb_range_heun', 'ibb'),
         # more than 16 generators (tiled grid, 20 machines with all controllers, RK4): wide stepper + wide adjoint
         ('grad_tiled20_rk4', 'tiled20')]


def loss_of(result, gold, names):
    idx = [names.index(str(c)) for c in gold['loss_channels']]
    out = result[:, :, idx]
    if 'target2d' in gold.files:                     # [T, C] target: the reference's golden trajectory (config 4)
        target = torch.from_numpy(gold['target2d'])[None].to(out.device)
    else:
        target = torch.from_numpy(gold['target'])[None, :, None].to(out.device) * torch.ones_like(out)
    return ((out - target) ** 2).mean(dim=(1, 2))


def assert_gradients_match(gold, leaves, to_numpy=lambda t: t.numpy()):
    """1e-6 relative (north star).  Per leaf against the largest gradient of the batch, and -- for rows that keep
    synchronism -- per ROW against that row's own gradient, so that one large row cannot hide a wrong small one."""
    names = [str(s) for s in gold['names']]
    dcols = [i for i, n in enumerate(names) if n.endswith('.delta')]
    stable = np.abs(gold['traj'][:, :, dcols]).max(axis=(1, 2)) < 10.0
    for (gname, attr), leaf in leaves.items():
        ref = gold['grad/gen|%s|%s' % (gname, attr)]
        mine = to_numpy(leaf.grad.real).reshape(-1)
        rel = np.abs(mine - ref) / (np.abs(ref).max() + 1e-300)
        assert rel.max() < 1e-6, (gname, attr, mine, ref)
        if ref.shape[0] == stable.shape[0]:
            big = stable & (np.abs(ref) > 1e-9 * np.abs(ref).max())
            row = np.abs(mine - ref)[big] / np.abs(ref)[big]
            assert (row < 1e-6).all(), (gname, attr, row.max())


@pytest.mark.parametrize('case,grid', CASES)
def test_adjoint_matches_reference_autograd(case, grid):
    sim, gold, leaves = helpers.sim_from_gold(case, grid, device='cpu', requires_grad=True)
    names = [str(s) for s in gold['names']]
    result = helpers.run_emulated_autograd(sim)
    loss = loss_of(result, gold, names)
    assert np.allclose(loss.detach().numpy(), gold['loss'], rtol=1e-9, atol=0)
    loss.sum().backward()
    assert_gradients_match(gold, leaves)
