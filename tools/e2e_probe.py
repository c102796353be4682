"""Where does the end-to-end time go?  (run on the GPU box)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench

P = 65536
batch = bench.parameter_batch(P)
sim = bench.build_sim(P, 'cuda', (0, P), batch)

def t(fn, n=3):
    torch.cuda.synchronize(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def prep():
    sim.reset(); return sim.prepare()
print('prepare            %.1f ms' % t(prep))
p = prep()
rec = torch.empty((10, bench.T_STEPS, P), dtype=torch.float64, device='cuda')
xf = torch.empty((10, 13, P), dtype=torch.float64, device='cuda')
print('launch (2 kernels) %.1f ms' % t(lambda: p.launch(rec, xf)))
host = torch.empty((10, bench.T_STEPS, P), dtype=torch.float64).pin_memory()
ms = t(lambda: host.copy_(rec, non_blocking=True))
print('D2H 10.5 GB        %.1f ms  (%.1f GB/s)' % (ms, rec.numel() * 8 / ms / 1e6))
def chunks():
    for k in range(8):
        for i in range(10):
            host[i, 250 * k:250 * (k + 1)].copy_(rec[i, 250 * k:250 * (k + 1)], non_blocking=True)
ms = t(chunks)
print('D2H in 80 pieces   %.1f ms  (%.1f GB/s)' % (ms, rec.numel() * 8 / ms / 1e6))
print('execute_ranges     %.1f ms' % t(lambda: p.execute_ranges([], host, 8)))
print('execute_ranges(2)  %.1f ms' % t(lambda: p.execute_ranges([], host, 2)))
def full():
    sim.reset(); sim.run(host_out=host, stream_chunks=8)
print('sim.run streamed   %.1f ms' % t(full))
def plain():
    sim.reset(); _, r = sim.run(); host.copy_(r.permute(2, 1, 0), non_blocking=True)
print('sim.run + copy     %.1f ms' % t(plain))
import cProfile, pstats
sim.reset()
pr = cProfile.Profile(); pr.enable(); sim.prepare(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
