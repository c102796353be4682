"""cProfile of prepare() in an identification-style loop (leaves require grad -> torch prologue)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from diffpssi_b200 import Pss
from tests import helpers
P = 16384
rng = np.random.RandomState(0)
sim = Pss(parallel_sims=P, sim_time=10.0, time_step=0.005, solver='heun', grid_data=helpers.load_grid('ibb'), verbose=False)
sim.add_sc_event(1, 1.05, 'Bus 1')
ov = {('Gen 1', 'h'): 3.5 * rng.uniform(0.5, 2, P), ('Gen 1', 'd'): rng.uniform(0, 2, P), ('Gen 1', 'x_d_t'): 0.3 * rng.uniform(0.5, 2, P)}
leaves = helpers.apply_overrides(sim, ov, requires_grad=True)
sim.set_record_channels(['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st'])
for it in range(3):
    sim.reset(); _, res = sim.run(); (res ** 2).mean().backward(); torch.cuda.synchronize()
sim.reset()
pr = cProfile.Profile(); pr.enable(); prep = sim.prepare(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(30)
