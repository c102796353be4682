import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench, cProfile, pstats
P = 65536
batch = bench.parameter_batch(P)
sim = bench.build_sim(P, 'cuda', (0, P), batch)
host_params = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in batch.items()}
host_out = torch.empty((10, bench.T_STEPS, P), dtype=torch.float64).pin_memory()
def e2e_step():
    bench.set_parameters(sim, host_params, (0, P), 'cuda')
    sim.reset()
    _, result = sim.run(host_out=host_out, stream_chunks=8)
    torch.cuda.synchronize()
e2e_step()
for i in range(3):
    t0 = time.perf_counter(); e2e_step(); print('e2e_step %.1f ms' % ((time.perf_counter() - t0) * 1e3), sim.timings)
pr = cProfile.Profile(); pr.enable(); e2e_step(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
