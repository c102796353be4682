"""GPU tests of initialize() on the device (csrc/dps_prologue.cu, diffpssi_b200/prologue_device.py):
the three prologue kernels against the differentiable torch prologue, which itself reproduces the reference's
initialize() bit for bit (tests/test_host_pipeline.py::test_prologue_matches_reference_initialize)."""
import numpy as np
import pytest
import torch

from diffpssi_b200 import Pss, grids, planning
from tests import helpers
from tests.test_cuda_parity import _random_overrides

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _build(grid, P, overrides=None, kernels=True):
    g = grids.ieee39() if grid == 'ieee39' else helpers.load_grid(grid)
    sim = Pss(parallel_sims=P, sim_time=0.5, time_step=0.005, solver='heun', grid_data=g, verbose=False, device='cuda')
    sim.prologue_kernels = kernels
    if overrides:
        helpers.apply_overrides(sim, overrides)
    return sim


def _overrides(grid, P, kind):
    rng = np.random.RandomState(5)
    if kind == 'uniform' or P == 1:
        return {}
    if grid == 'ieee39':
        ov = {}
        for g, (h, x) in grids.GENERATOR_NOMINALS['ieee39'].items():
            ov[(g, 'h')] = h * rng.uniform(0.5, 2, P)
            ov[(g, 'd')] = rng.uniform(0, 2, P)
            ov[(g, 'x_d_t')] = x * rng.uniform(0.8, 1.25, P)
        return ov
    if grid.startswith('tiled'):                # 100 buses: torch LF / inverse + the model-initialisation kernel
        return {('M0', 'h'): 6.5 * rng.uniform(0.8, 1.25, P), ('M1', 'd'): rng.uniform(0, 2, P),
                ('M2', 'x_d_t'): 0.3 * rng.uniform(0.8, 1.25, P), ('M3', 'exciter.pt1_lim.gain_pt1'): 100.0 * rng.uniform(0.8, 1.25, P)}
    ov = _random_overrides(grid, P, rng, 0.8, 1.25)
    if kind == 'network':                       # a per-trajectory network: x_d'' enters Y_dyn, a controller constant too
        first = sorted({k[0] for k in ov})[0]
        nominal = {'k2a': 0.25, 'ieee9': 0.1, 'ibb': 0.2, 'ibb_wc': 0.2}[grid]
        ov[(first, 'x_d_st')] = nominal * rng.uniform(0.9, 1.1, P)
    return ov


def _compare(a, b, what, tol=TOL):
    a, b = a.detach().cpu().numpy(), b.detach().cpu().numpy()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b).max() if a.size else 0.0
    assert err < tol, (what, err)


@pytest.mark.parametrize('grid,P,kind', [
    ('ibb', 1, 'uniform'), ('ibb', 33, 'persim'), ('ibb_wc', 70, 'persim'), ('ibb_wc', 40, 'network'),
    ('ieee9', 1, 'uniform'), ('ieee9', 97, 'persim'), ('k2a', 200, 'persim'), ('k2a', 64, 'network'),
    ('ieee39', 3, 'uniform'), ('ieee39', 512, 'persim'), ('tiled20', 24, 'persim'),
])
def test_kernel_prologue_equals_torch_prologue(grid, P, kind):
    ov = _overrides(grid, P, kind)
    out = {}
    for kernels in (True, False):
        sim = _build(grid, P, ov, kernels)
        init = sim.initialize()
        assert sim.last_prologue == ('kernels' if kernels else 'torch')
        x0 = planning.canonical_x0(init, P)
        upar, vpar, slot_vary = planning.pack_slots(init, P, 'cuda')
        volt = torch.cat([b.voltage for b in sim.busses], dim=1)
        loads = [m.y_load for bus in sim.busses for m in bus.models if hasattr(m, 'y_load')]
        out[kernels] = (init, x0, upar, vpar, slot_vary, volt, loads)
    (ik, x0k, uk, vk, svk, voltk, lk), (it, x0t, ut, vt, svt, voltt, lt) = out[True], out[False]
    assert ik.lf_iters == it.lf_iters and ik.ctrl == it.ctrl and ik.gen_bus == it.gen_bus
    _compare(ik.bus_v, it.bus_v, 'load-flow voltage')
    _compare(ik.s_calc, it.s_calc, 'load-flow power', 1e-11)
    _compare(ik.Z0, it.Z0, 'Z0', 1e-11)
    _compare(ik.Ydiag, it.Ydiag, 'diag(Y_dyn)', 1e-10)
    _compare(x0k, x0t, 'x0')
    # the same constants are classified per-trajectory by both paths, and carry the same values
    assert svk == svt
    live = torch.tensor([s < 0 for s in svk], device='cuda').reshape(uk.shape)
    _compare(torch.where(live, uk, torch.zeros_like(uk)), torch.where(live, ut, torch.zeros_like(ut)), 'upar', 1e-11)
    assert (vk is None) == (vt is None)
    if vk is not None:
        _compare(vk, vt, 'vpar', 1e-11)
    _compare(voltk, voltt, 'initial bus voltages', 1e-11)
    assert len(lk) == len(lt)
    for a, b in zip(lk, lt):
        _compare(a, b, 'y_load', 1e-11)


@pytest.mark.parametrize('grid,P,event', [('k2a', 48, (0.1, 0.2, 'B8')), ('ieee39', 96, (0.1, 0.2, 'B16'))])
def test_run_with_kernel_prologue_matches_torch_prologue(grid, P, event):
    """Whole trajectories (every state of every generator) with either prologue, incl. a rerun after reset() that hits the
    memoised load flow."""
    ov = _overrides(grid, P, 'persim')
    res = {}
    for kernels in (True, False):
        sim = _build(grid, P, ov, kernels)
        sim.add_sc_event(*event)
        init = sim.initialize()
        sim.set_record_channels(['%s.%s' % (g.name, f) for g in init.gens for f in ('omega', 'delta', 'e_q_st', 'p_e', 'e_fd', 'p_m')])
        sim.reset()                 # (initialize() left the initial bus voltages behind; start both runs from the flat start)
        _, r1 = sim.run()
        sim.reset()
        _, r2 = sim.run()
        assert torch.equal(r1, r2)
        res[kernels] = r1.cpu().numpy()
    assert np.abs(res[True] - res[False]).max() < 1e-10


def test_kernel_prologue_steps_aside_for_gradients():
    """A leaf that requires grad keeps the differentiable torch prologue (the autograd contract of
    synchronous_machine.py:294-314 is the torch path's)."""
    sim = _build('ibb', 8)
    helpers.apply_overrides(sim, {('Gen 1', 'h'): np.full(8, 3.5)}, requires_grad=True)
    sim.initialize()
    assert sim.last_prologue == 'torch'
    with torch.no_grad():
        sim.reset()
        sim.initialize()
        assert sim.last_prologue == 'kernels'
