"""Timings of the other BASELINE.json configurations (parity-test cases, not the bench line): run on the GPU box.

    python tools/bench_configs.py [c2] [c4] [c5]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from diffpssi_b200 import Pss, grids
from tests import helpers


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def loop_only(sim, n=5):
    """Kernel-only time of the prepared run (as bench.py's `value`)."""
    prep = sim.prepare()
    topo = prep.topo
    P = sim.parallel_sims
    rec = torch.empty((len(topo['rec']), topo['n_steps'], P), dtype=torch.float64, device='cuda')
    xf = torch.empty((topo['n_gen'], 13, P), dtype=torch.float64, device='cuda')
    return timed(lambda: prep.launch(rec, xf), n), prep.plan.info()


def c2():
    P, T = 4096, 2000
    rng = np.random.RandomState(0)
    sim = Pss(parallel_sims=P, sim_time=10, time_step=0.005, solver='heun', grid_data=helpers.load_grid('k2a'), verbose=False)
    sim.add_sc_event(1, 1.1, 'B1')
    ov = {}
    for g, (h, x) in grids.GENERATOR_NOMINALS['k2a'].items():
        ov[(g, 'h')] = h * rng.uniform(0.5, 2, P); ov[(g, 'd')] = rng.uniform(0, 2, P); ov[(g, 'x_d_t')] = x * rng.uniform(0.5, 2, P)
    helpers.apply_overrides(sim, ov)
    sim.set_record_channels(['G%d.%s' % (i, f) for f in ('delta', 'omega', 'e_q_st', 'e_d_st') for i in (1, 2, 3, 4)])
    dt, info = loop_only(sim)
    return {'config': 'C2 K2A heun P=4096 T=2000', 'loop_ms': dt * 1e3, 'sim_steps_per_s': P * T / dt, 'kernel': info}


def c4():
    P, T = 16384, 2000
    rng = np.random.RandomState(0)
    out = {}
    for solver in ('heun',):
        sim = Pss(parallel_sims=P, sim_time=10, time_step=0.005, solver=solver, grid_data=helpers.load_grid('ibb'), verbose=False)
        sim.add_sc_event(1, 1.05, 'Bus 1')
        ov = {('Gen 1', 'h'): 3.5 * rng.uniform(0.5, 2, P), ('Gen 1', 'd'): rng.uniform(0, 2, P),
              ('Gen 1', 'x_d_t'): 0.3 * rng.uniform(0.5, 2, P)}
        leaves = helpers.apply_overrides(sim, ov, requires_grad=True)
        sim.set_record_channels(['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st'])
        target = torch.from_numpy(np.load(os.path.join(helpers.GOLD, 'ref_ibb_model.npy'))).cuda()

        def fwd_bwd():
            for l in leaves.values():
                l.grad = None
            sim.reset()
            _, res = sim.run()
            loss = ((res - target[None]) ** 2).mean(dim=(1, 2))
            loss.sum().backward()

        def fwd_only():
            sim.reset()
            with torch.no_grad():
                sim.run()

        t_fb = timed(fwd_bwd)
        t_f = timed(fwd_only)
        out = {'config': 'C4 IBB heun P=16384 T=2000 forward+adjoint (3 leaves)', 'fwd_bwd_ms': t_fb * 1e3,
               'fwd_only_ms': t_f * 1e3, 'fwd_bwd_sim_steps_per_s': P * T / t_fb,
               'grad_finite_fraction': float(torch.isfinite(list(leaves.values())[0].grad.real).double().mean())}
    return out


def c5():
    P, T = 8192, 2000
    rng = np.random.RandomState(0)
    sim = Pss(parallel_sims=P, sim_time=T * 0.005, time_step=0.005, solver='rk4', grid_data=grids.tiled(100), verbose=False)
    sim.add_sc_event(1.0, 1.1, 'A3')
    ov = {}
    for c in range(100):
        ov[('M%d' % c, 'h')] = (6.5 if c % 2 == 0 else 6.175) * rng.uniform(0.5, 2, P)
        ov[('M%d' % c, 'd')] = rng.uniform(0, 2, P)
        ov[('M%d' % c, 'x_d_t')] = 0.3 * rng.uniform(0.5, 2, P)
    helpers.apply_overrides(sim, ov)
    sim.set_record_channels(['M%d.omega' % c for c in range(100)])
    dt, info = loop_only(sim, n=2)
    flops = 4 * 100 * (106 + 800) + 14 * 1300
    return {'config': 'C5 tiled 500-bus / 100-gen rk4 P=8192 T=2000', 'loop_ms': dt * 1e3, 'sim_steps_per_s': P * T / dt,
            'algorithmic_tflops': flops * P * T / dt / 1e12, 'kernel': info}


if __name__ == '__main__':
    which = sys.argv[1:] or ['c2', 'c4', 'c5']
    for w in which:
        print(json.dumps(globals()[w]()))
