"""cProfile of a cold and a warm kernel-path initialize() (run on the GPU box)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from diffpssi_b200 import Pss, grids
from tests import helpers
P = 8192

def build():
    sim = Pss(parallel_sims=P, sim_time=10.0, time_step=0.005, solver='heun', grid_data=grids.ieee39(), verbose=False, device='cuda')
    rng = np.random.RandomState(0)
    ov = {}
    for g, (h, x) in grids.GENERATOR_NOMINALS['ieee39'].items():
        ov[(g, 'h')] = h * rng.uniform(0.5, 2, P); ov[(g, 'd')] = rng.uniform(0, 2, P); ov[(g, 'x_d_t')] = x * rng.uniform(0.5, 2, P)
    helpers.apply_overrides(sim, ov)
    sim.add_sc_event(1.0, 1.1, 'B16')
    sim.set_record_channels(['G%d.omega' % i for i in range(1, 11)])
    return sim

build().prepare(); torch.cuda.synchronize()
sim = build()
pr = cProfile.Profile(); pr.enable(); sim.initialize(); torch.cuda.synchronize(); pr.disable()
print('==== cold initialize()'); pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
sim.reset()
pr = cProfile.Profile(); pr.enable(); sim.prepare(); torch.cuda.synchronize(); pr.disable()
print('==== warm prepare()'); pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
