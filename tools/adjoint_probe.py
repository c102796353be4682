"""Where does an identification step spend its time?  (run on the GPU box)
    [CKPT=<stride>] [SOLVER=rk4] [CELLS=100] python tools/adjoint_probe.py [ibb|k2a|ieee39|tiled] [P]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from diffpssi_b200 import Pss, grids, adjoint
from tests import helpers

which = sys.argv[1] if len(sys.argv) > 1 else 'ibb'
P = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
T = int(os.environ.get("STEPS", 2000))
rng = np.random.RandomState(0)
if which == 'ibb':
    grid, ev = helpers.load_grid('ibb'), (1, 1.05, 'Bus 1')
    ov = {('Gen 1', 'h'): 3.5 * rng.uniform(0.5, 2, P), ('Gen 1', 'd'): rng.uniform(0, 2, P), ('Gen 1', 'x_d_t'): 0.3 * rng.uniform(0.5, 2, P)}
    chans = ['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st']
elif which == 'k2a':
    grid, ev = helpers.load_grid('k2a'), (1, 1.1, 'B1')
    ov = {}
    for g, (h, x) in grids.GENERATOR_NOMINALS['k2a'].items():
        ov[(g, 'h')] = h * rng.uniform(0.8, 1.25, P); ov[(g, 'd')] = rng.uniform(0, 2, P)
    chans = ['G%d.omega' % i for i in (1, 2, 3, 4)]
elif which == 'tiled':                           # BASELINE config 5's grid (500 buses / 100 generators): the wide adjoint
    n_cells = int(os.environ.get('CELLS', 100))
    grid, ev = grids.tiled(n_cells), (1.0, 1.1, 'A3')
    ov = {}
    for c in range(n_cells):
        ov[('M%d' % c, 'h')] = (6.5 if c % 2 == 0 else 6.175) * rng.uniform(0.8, 1.25, P)
        ov[('M%d' % c, 'd')] = rng.uniform(0, 2, P)
    chans = ['M%d.omega' % c for c in range(n_cells)]
else:
    grid, ev = grids.ieee39(), (1.0, 1.1, 'B16')
    ov = {}
    for g, (h, x) in grids.GENERATOR_NOMINALS['ieee39'].items():
        ov[(g, 'h')] = h * rng.uniform(0.8, 1.25, P); ov[(g, 'd')] = rng.uniform(0, 2, P)
    chans = ['G%d.omega' % i for i in range(1, 11)]
sim = Pss(parallel_sims=P, sim_time=T * 0.005, time_step=0.005, solver=os.environ.get('SOLVER', 'heun'), grid_data=grid, verbose=False)
sim.add_sc_event(*ev)
leaves = helpers.apply_overrides(sim, ov, requires_grad=True)
sim.ckpt_stride = int(os.environ['CKPT']) if 'CKPT' in os.environ else None      # None: Pss picks (1 when it fits)
sim.set_record_channels(chans)

def sync(): torch.cuda.synchronize()
times = {}
orig = adjoint.run_backward
def timed_backward(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = orig(*a, **k); e1.record(); sync()
    times['dps_backward_ms'] = e0.elapsed_time(e1)
    return out
adjoint.run_backward = timed_backward
for it in range(3):
    for l in leaves.values(): l.grad = None
    sim.reset(); sync(); t0 = time.perf_counter()
    _, res = sim.run(); sync(); t1 = time.perf_counter()
    loss = (res ** 2).mean(dim=(1, 2)).sum(); sync(); t2 = time.perf_counter()
    loss.backward(); sync(); t3 = time.perf_counter()
print('ckpt_stride %s -> %d ' % (sim.ckpt_stride, sim.last_plan.topo['ckpt_stride']), end='')
print('%s P=%d T=%d: run() %.1f ms (prologue %.1f), loss %.1f ms, backward() %.1f ms of which dps_backward (+alloc/scatter) %.1f ms'
      % (which, P, T, (t1 - t0) * 1e3, sim.timings['prologue_s'] * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, times['dps_backward_ms']))
with torch.no_grad():
    sim.reset(); sync(); t0 = time.perf_counter(); sim.run(); sync()
    print('   forward without grad: %.1f ms' % ((time.perf_counter() - t0) * 1e3))
