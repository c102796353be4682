"""Prologue cost on the GPU box: kernel path (csrc/dps_prologue.cu) vs the differentiable torch path, cold (first call of a
fresh simulation object, nothing memoised) and warm (after reset(), the memoised load flow / inverse reused).
    python tools/prologue_probe.py [P]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from diffpssi_b200 import Pss, grids
from tests import helpers

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8192


def build(kernels, pdev=None):
    sim = Pss(parallel_sims=P, sim_time=10.0, time_step=0.005, solver='heun', grid_data=grids.ieee39(), verbose=False,
              device='cuda', prologue_device=pdev)
    sim.prologue_kernels = kernels
    rng = np.random.RandomState(0)
    ov = {}
    for g, (h, x) in grids.GENERATOR_NOMINALS['ieee39'].items():
        ov[(g, 'h')] = h * rng.uniform(0.5, 2, P)
        ov[(g, 'd')] = rng.uniform(0, 2, P)
        ov[(g, 'x_d_t')] = x * rng.uniform(0.5, 2, P)
    helpers.apply_overrides(sim, ov)
    sim.add_sc_event(1.0, 1.1, 'B16')
    sim.set_record_channels(['G%d.omega' % i for i in range(1, 11)])
    return sim, ov


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


build(True)[0].prepare()                                                # CUDA context, library, allocator warm-up
for label, kernels, pdev in (('kernels', True, None), ('torch on the GPU', False, None), ('torch on the host', False, 'cpu')):
    cold = []
    for rep in range(3):
        sim, ov = build(kernels, pdev)
        cold.append((timed(sim.initialize), sim))
    sim = cold[-1][1]
    t_prep = []
    for rep in range(5):
        sim.reset()
        t_prep.append(timed(sim.prepare))
    # an identification-style rerun: new per-trajectory tensors, everything else memoised
    t_new = []
    for rep in range(5):
        sim.reset()
        helpers.apply_overrides(sim, {k: v * 1.0001 for k, v in ov.items()})
        t_new.append(timed(sim.prepare))
    print('%-18s IEEE-39 P=%d: initialize() cold %.2f / %.2f / %.2f ms;  prepare() after reset %.2f ms;  prepare() with new H, D, X_d\' tensors %.2f ms  [%s]'
          % (label, P, cold[0][0], cold[1][0], cold[2][0], min(t_prep), min(t_new), sim.last_prologue))
