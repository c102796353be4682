"""GPU tests of the device-side identification step (csrc/dps_optim.cu, diffpssi_b200/optimization.py) against known
answers produced by the reference's own classes (oracle/gen_golden.py: optimizer_cases):
CustomBFGSREALOptimizer (optimization_lib/optimizers.py) and PowerSystemOptimization.run() (ps_optimization.py)."""
import numpy as np
import pytest
import torch

from diffpssi_b200 import Pss, CustomBFGSREALOptimizer, PowerSystemOptimization, engine
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('tag', ['', '_n5', '_n1'])
def test_bfgs_kernel_matches_reference_optimizer(tag):
    """dps_bfgs_step vs the reference optimiser on seeded gradients: parameter trajectory over 7 steps incl. two step-size
    decays, and the final inverse Hessian.  Floating point: 1e-10 relative (the kernel and torch.matmul sum the small
    products in a different order)."""
    gold = helpers.load_gold('bfgs_reference' + tag)
    x0, grads = gold['x0'], gold['grads']
    n, P = x0.shape
    params = [torch.tensor(x0[i].reshape(P, 1), dtype=torch.complex128, device='cuda', requires_grad=True) for i in range(n)]
    opt = CustomBFGSREALOptimizer(params, max_step=float(gold['max_step']), decay=float(gold['decay']), verbose=False)
    for k in range(grads.shape[0]):
        for i, p in enumerate(params):
            g = torch.tensor(grads[k, i].reshape(P, 1), device='cuda')
            p.grad = torch.complex(g, torch.full_like(g, 0.123 * (k + 1)))          # imaginary part: garbage, ignored
        opt.step()
        if k in gold['decrease_after']:
            opt.decrease_step_size()
        mine = np.stack([p.data.real.cpu().numpy().reshape(-1) for p in params])
        assert np.allclose(mine, gold['traj'][k], rtol=1e-10, atol=0), (k, np.abs(mine / gold['traj'][k] - 1).max())
        assert abs(opt.max_step - gold['max_step_after'][k]) < 1e-15
        assert all(float(p.data.imag.abs().max()) == 0.0 for p in params)
    hesse = opt.hesse.permute(2, 0, 1).cpu().numpy()                                # [P,n,n] like the reference
    ref = gold['hesse']
    assert np.allclose(hesse, ref, rtol=1e-8, atol=1e-12 * np.abs(ref).max())


def test_fused_loss_matches_torch():
    """dps_ident_loss (normalised mean absolute error + its gradient + batch arg-min) vs the same expression in torch
    with autograd (ps_optimization.py:87-101, 194-195, 202, 212), shared and per-trajectory targets, NaN rows."""
    rng = np.random.RandomState(0)
    P, T, C = 777, 300, 3
    res = torch.tensor(rng.normal(0, 1, (P, T, C)), device='cuda')
    res[5, 17, 1] = float('nan')
    for Pt in (1, P):
        target = torch.tensor(rng.normal(0, 1, (Pt, T, C)), device='cuda')
        target[0, 3, 0] = float(res[0, 3, 0])                                       # an exact tie: sign(0) = 0
        vmin = torch.min(target, dim=1)[0].unsqueeze(1)
        vrange = torch.max(target, dim=1)[0].unsqueeze(1) - vmin
        r = res.clone().requires_grad_(True)
        want = torch.mean(torch.sum(torch.abs((target - vmin) / vrange - (r - vmin) / vrange), dim=2), dim=1)
        want.sum().backward()
        state = torch.zeros(4, dtype=torch.float64, device='cuda')
        loss, g, _ = engine.ident_loss(res.permute(2, 1, 0).contiguous(), target.permute(2, 1, 0).contiguous(),
                                       vmin[:, 0, :].permute(1, 0).contiguous(), vrange[:, 0, :].permute(1, 0).contiguous(),
                                       state)
        ok = torch.isfinite(want)
        assert torch.equal(torch.isfinite(loss), ok)
        assert float((loss[ok] - want[ok]).abs().max()) < 1e-13
        gw = r.grad.permute(2, 1, 0)
        fin = torch.isfinite(gw)
        assert float((g[fin] - gw[fin]).abs().max()) < 1e-16
        m, i = torch.nan_to_num(want.detach(), 100000).min(dim=0)
        assert int(state[3]) == int(i) and abs(float(state[2]) - float(m)) < 1e-13


def _ident_problem(cuda_graph, steps):
    gold = helpers.load_gold('opt_ibb_reference')
    keys = [str(k).split('|') for k in gold['keys']]
    P = gold['traj'].shape[2]
    sim = Pss(parallel_sims=P, sim_time=3.0, time_step=0.005, solver='heun', grid_data=helpers.load_grid('ibb'), verbose=False)
    sim.add_sc_event(1, 1.05, 'Bus 1')
    ov = {(k[1], k[2]): gold['override/' + '|'.join(k)] for k in keys}
    leaves = helpers.apply_overrides(sim, ov, requires_grad=True)
    sim.set_record_function(lambda s: [s.busses[1].models[0].omega.real, s.busses[1].models[0].e_q_st.real,
                                       s.busses[1].models[0].e_d_st.real])
    params = [leaves[(k[1], k[2])] for k in keys]
    target = torch.tensor(gold['target'], device='cuda') * torch.ones((P,) + gold['target'].shape, dtype=torch.float64, device='cuda')
    opt = PowerSystemOptimization(sim, target, params_optimizable=params, param_names=['|'.join(k) for k in keys],
                                  decay=float(gold['decay']), max_step=float(gold['max_step']), verbose=False,
                                  cuda_graph=cuda_graph)
    traj = []
    orig = opt._optimizer_step

    def spy(k):
        orig(k)
        traj.append(torch.stack([p.data.real.reshape(-1).clone() for p in params]))
    opt._optimizer_step = spy
    opt.run(max_steps=steps)
    return gold, opt, params, traj


def test_identification_loop_matches_reference():
    """PowerSystemOptimization.run() on the IBB model (3 leaves, 6 trajectories, 4 steps): the parameters after every
    optimiser step against the reference's own loop (gradients through dps_backward, loss / BFGS / step control on the
    device).  1e-6 relative, the north-star gradient tolerance."""
    gold, opt, params, traj = _ident_problem(False, gold_steps := 4)
    assert len(traj) == gold_steps
    for k in range(gold_steps):
        mine = traj[k].cpu().numpy()
        assert np.allclose(mine, gold['traj'][k], rtol=1e-6, atol=0), (k, np.abs(mine / gold['traj'][k] - 1).max())
    assert abs(opt.optimizer.max_step - float(gold['final_max_step'])) < 1e-15


def test_graphed_identification_equals_eager():
    """cuda_graph=True: from the fourth step on one captured CUDA graph (prologue, kernels, loss, adjoint, optimiser) is
    replayed; the parameter trajectory must equal the eager loop bit for bit."""
    _, _, _, eager = _ident_problem(False, 7)
    _, opt, _, graphed = _ident_problem(True, 7)
    assert opt._graph is not None
    final = torch.stack([p.data.real.reshape(-1) for p in opt.params])
    assert torch.equal(torch.stack(eager[:3]), torch.stack(graphed[:3]))
    assert torch.equal(final, eager[-1])
