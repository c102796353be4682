"""CI-size pass over every kernel family for ``compute-sanitizer --tool racecheck`` (shared-memory hazards: the per-warp
exchange lines behind __syncwarp, the Z / constant tables, the MMA operand buffers of the wide stepper, the work-queue
ticket) -- run under gpurun:

    compute-sanitizer --tool racecheck python tools/racecheck.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from diffpssi_b200 import Pss, grids
from tests import helpers


def k2a(P, T, solver, grad=False):
    sim = Pss(parallel_sims=P, sim_time=T * 0.005, time_step=0.005, solver=solver, grid_data=helpers.load_grid('k2a'), verbose=False)
    sim.add_sc_event(0.05, 0.1, 'B8')
    rng = np.random.RandomState(0)
    leaves = helpers.apply_overrides(sim, {('G1', 'h'): 6.5 * rng.uniform(0.9, 1.1, P), ('G2', 'd'): rng.uniform(0, 2, P)},
                                     requires_grad=grad)
    sim.set_record_channels(['G%d.%s' % (i, f) for f in ('omega', 'p_e') for i in (1, 2, 3, 4)] + ['bus:B8.v_re'])
    _, res = sim.run()
    if grad:
        (res ** 2).mean(dim=(1, 2)).sum().backward()
    torch.cuda.synchronize()
    return float(res.abs().max())


def ieee39(P, T, pair, queue_slots=None):
    os.environ['DPS_PAIR_MIN_SIMS'] = '0' if pair else '999999999'
    if queue_slots:
        os.environ['DPS_QUEUE_SLOTS'] = str(queue_slots)
    else:
        os.environ.pop('DPS_QUEUE_SLOTS', None)
    sim = Pss(parallel_sims=P, sim_time=T * 0.005, time_step=0.005, solver='heun', grid_data=grids.ieee39(), verbose=False)
    sim.add_sc_event(0.05, 0.1, 'B16')
    rng = np.random.RandomState(1)
    helpers.apply_overrides(sim, {('G%d' % i, 'h'): 4.0 * rng.uniform(0.8, 1.2, P) for i in range(1, 11)})
    sim.set_record_channels(['G%d.omega' % i for i in range(1, 11)])
    _, res = sim.run()
    torch.cuda.synchronize()
    return float(res.abs().max())


def tiled(P, T):
    sim = Pss(parallel_sims=P, sim_time=T * 0.005, time_step=0.005, solver='rk4', grid_data=grids.tiled(20), verbose=False)
    sim.add_sc_event(0.02, 0.04, 'A3')
    sim.set_record_channels(['M%d.omega' % c for c in range(20)] + ['bus:A3.v_re'])
    _, res = sim.run()
    torch.cuda.synchronize()
    return float(res.abs().max())


if __name__ == '__main__':
    print('step kernel (K2A heun/rk4):', k2a(40, 30, 'heun'), k2a(9, 20, 'rk4'))
    print('adjoint kernel (K2A heun, stride 32 and 1):', k2a(40, 40, 'heun', grad=True))
    print('pair kernel:', ieee39(77, 30, True))
    print('pair kernel, work queue:', ieee39(150, 96, True, queue_slots=2))
    print('step kernel, work queue:', ieee39(150, 96, False, queue_slots=3))
    print('wide kernel (tiled 20 cells, rk4):', tiled(9, 12))
    print('racecheck pass complete')
