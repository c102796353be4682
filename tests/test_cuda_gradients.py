"""GPU gradient tests: dps_backward (hand-written discrete adjoint, csrc/dps_backward.cu) behind torch.autograd,
against the reference's own autograd gradients (tests/golden/grad_*.npz; generated from the reference's torch backend
and cross-checked by finite differences there).  Tolerance 1e-6 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from diffpssi_b200 import Pss
from tests import helpers
from tests.test_gradients_host import CASES, loss_of, assert_gradients_match

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case,grid', CASES)
@pytest.mark.parametrize('stride', [32, 7, 1, None])
def test_adjoint_matches_reference_autograd(case, grid, stride):
    sim, gold, leaves = helpers.sim_from_gold(case, grid, device='cuda', requires_grad=True)
    sim.ckpt_stride = stride       # 7 does not divide the step count (ragged last segment); 1 = every step kept, no
                                   # re-integration in the adjoint; None = Pss chooses (1 for these sizes)
    names = [str(s) for s in gold['names']]
    _, result = sim.run()
    loss = loss_of(result.cpu(), gold, names)
    assert np.allclose(loss.detach().numpy(), gold['loss'], rtol=1e-9, atol=0)
    loss.sum().backward()
    assert_gradients_match(gold, leaves, to_numpy=lambda t: t.cpu().numpy())


def test_gradients_tiled_batch_and_shared_parameter():
    """Batch of 70 (ragged), the same 4 golden rows tiled: every copy gets the golden per-sim gradient; a *shared*
    (uniform) parameter tensor gets the sum over the batch."""
    tile = 18
    sim, gold, leaves = helpers.sim_from_gold('grad_ibb_heun', 'ibb', device='cuda', requires_grad=True, tile=tile)
    P = sim.parallel_sims
    gen = sim.get_generator_by_name('Gen 1')
    shared = torch.tensor([[0.07 + 0j]], dtype=torch.complex128, device='cuda', requires_grad=True)
    gen.t_q0_st = shared                           # [1,1]: uniform slot with a gradient
    names = [str(s) for s in gold['names']]
    _, result = sim.run()
    idx = [names.index(str(c)) for c in gold['loss_channels']]
    out = result[:, :, idx]
    target = torch.from_numpy(gold['target']).to(out.device)[None, :, None] * torch.ones_like(out)
    loss = ((out - target) ** 2).mean(dim=(1, 2))
    loss.sum().backward()
    for (gname, attr), leaf in leaves.items():
        ref = np.tile(gold['grad/gen|%s|%s' % (gname, attr)], tile)
        mine = leaf.grad.real.cpu().numpy().reshape(-1)
        assert (np.abs(mine - ref) / np.abs(ref).max()).max() < 1e-6
    assert shared.grad is not None and abs(float(shared.grad.real)) > 0
    # finite-difference check of the shared-parameter gradient through the whole public API
    def total_loss(v):
        s2, _, _ = helpers.sim_from_gold('grad_ibb_heun', 'ibb', device='cuda', tile=tile)
        s2.get_generator_by_name('Gen 1').t_q0_st = torch.tensor([[v + 0j]], dtype=torch.complex128, device='cuda')
        _, r = s2.run()
        o = r[:, :, idx]
        return float(((o - target) ** 2).mean(dim=(1, 2)).sum())
    h = 1e-6
    fd = (total_loss(0.07 + h) - total_loss(0.07 - h)) / (2 * h)
    assert abs(fd - float(shared.grad.real)) / abs(fd) < 1e-5


def test_shared_network_parameter_gradient():
    """A *uniform* x_d'' leaf ([1,1]) keeps the network shared by the batch: dL/dZsub is summed over the trajectories
    and chained through the fault re-reduction and inv(Y_dyn).  Checked by central finite differences through the
    public API (the reference itself cannot differentiate this case: its in-place fault switching breaks autograd)."""
    tile = 9
    def build(v, grad):
        sim, gold, _ = helpers.sim_from_gold('grad_ibb_heun', 'ibb', device='cuda', tile=tile)
        leaf = torch.tensor([[v + 0j]], dtype=torch.complex128, device='cuda', requires_grad=grad)
        sim.get_generator_by_name('Gen 1').x_d_st = leaf
        return sim, gold, leaf
    def loss_of_run(sim, gold):
        names = [str(s) for s in gold['names']]
        idx = [names.index(str(c)) for c in gold['loss_channels']]
        _, r = sim.run()
        o = r[:, :, idx]
        target = torch.from_numpy(gold['target']).to(o.device)[None, :, None] * torch.ones_like(o)
        return ((o - target) ** 2).mean(dim=(1, 2)).sum()
    sim, gold, leaf = build(0.23, True)
    loss = loss_of_run(sim, gold)
    assert sim.last_plan.topo['z_per_sim'] == 0 and sim.last_plan.topo['want_z']
    loss.backward()
    g = float(leaf.grad.real)
    h = 1e-6
    lp = float(loss_of_run(*build(0.23 + h, False)[:2]))
    lm = float(loss_of_run(*build(0.23 - h, False)[:2]))
    fd = (lp - lm) / (2 * h)
    assert abs(fd) > 0 and abs(fd - g) / abs(fd) < 1e-5, (fd, g)


def test_identification_loop_reduces_the_loss():
    """A miniature of the reference's parameter identification (examples/models/ibb_model/ibb_opt.py,
    optimization_lib/ps_optimization.py:178-212): the target trajectory comes from the nominal H, every simulation of the
    batch starts from a wrong H, and a few gradient steps on the per-sim H (gradients from dps_backward through the
    prologue graph) must pull every trajectory towards the target."""
    P = 16
    grid = helpers.load_grid('ibb')

    def run(h_values, leaf=None):
        sim = Pss(parallel_sims=P, sim_time=2.0, time_step=0.005, solver='heun', grid_data=grid, verbose=False)
        sim.add_sc_event(1.0, 1.05, 'Bus 1')
        gen = sim.get_generator_by_name('Gen 1')
        gen.h = leaf if leaf is not None else torch.tensor(h_values.reshape(P, 1), dtype=torch.complex128, device='cuda')
        sim.set_record_channels(['Gen 1.omega'])
        _, res = sim.run()
        return res

    with torch.no_grad():
        target = run(np.full(P, 3.5))[:1]                                  # nominal H = 3.5 s
    h = torch.tensor(np.linspace(2.0, 6.0, P).reshape(P, 1), dtype=torch.complex128, device='cuda', requires_grad=True)
    opt = torch.optim.Adam([h], lr=0.15)
    losses = []
    for _ in range(12):
        opt.zero_grad()
        loss = ((run(None, leaf=h) - target) ** 2).mean(dim=(1, 2))
        losses.append(loss.detach().clone())
        loss.sum().backward()
        opt.step()
    first, last = losses[0], losses[-1]
    off = (h.detach().real.reshape(-1) - 3.5).abs() > 0.3                  # sims that did not start next to the optimum
    assert bool((last[off] < 0.5 * first[off]).all()), (first, last)
    assert float((h.detach().real.reshape(-1) - 3.5).abs().mean()) < float(np.abs(np.linspace(2.0, 6.0, P) - 3.5).mean())


@pytest.mark.parametrize('stride', [32, 5, 1])
def test_adjoint_work_queue_matches_static_launch(monkeypatch, stride):
    """dps_adjoint_kernel as (block group, chunk of checkpoint segments) work items (forced at a CI batch size through
    DPS_QUEUE_SLOTS): co-state and slot-gradient accumulators travel between the chunks through g_x0 / gslot; the gradients
    must be bit-identical to the one-block-per-group launch."""
    grads = {}
    for slots in (None, '2'):
        if slots:
            monkeypatch.setenv('DPS_QUEUE_SLOTS', slots)
        else:
            monkeypatch.delenv('DPS_QUEUE_SLOTS', raising=False)
        sim, gold, leaves = helpers.sim_from_gold('grad_k2a_heun', 'k2a', device='cuda', requires_grad=True, tile=40)
        sim.ckpt_stride = stride
        names = [str(s) for s in gold['names']]
        _, result = sim.run()
        loss = loss_of(result, gold, names)
        loss.sum().backward()
        grads[slots] = {k: v.grad.real.clone() for k, v in leaves.items()}
    for k in grads[None]:
        assert torch.equal(grads[None][k], grads['2'][k]), k
        ref = np.tile(gold['grad/gen|%s|%s' % k], 40)
        mine = grads['2'][k].cpu().numpy().reshape(-1)
        assert (np.abs(mine - ref) / np.abs(ref).max()).max() < 1e-6


@pytest.mark.parametrize('stride', [None, 7])
def test_adjoint_through_parameter_events(stride):
    """Gradients THROUGH ParamEvents (static_models.py:52-95): the run is cut at the event steps, every range is its own
    autograd node and the co-state is handed from one range's adjoint to the previous one's (dps_backward: grad_x_final).
    Golden: the reference's own autograd on the same run (D overwritten at t = 1.0, H at t = 1.5; FD-checked there)."""
    sim, gold, leaves = helpers.sim_from_gold('grad_ibb_param_event_heun', 'ibb', device='cuda', requires_grad=True)
    sim.ckpt_stride = stride
    names = [str(s) for s in gold['names']]
    _, result = sim.run()
    loss = loss_of(result.cpu(), gold, names)
    assert np.allclose(loss.detach().numpy(), gold['loss'], rtol=1e-9, atol=0)
    loss.sum().backward()
    assert_gradients_match(gold, leaves, to_numpy=lambda t: t.cpu().numpy())
    # the leaf of D stops collecting gradient at its event: compare with an event-free run of the same leaves
    assert np.abs(leaves[('Gen 1', 'd')].grad.real.cpu().numpy()).max() > 0
