#!/usr/bin/env python
"""Times the UNMODIFIED reference (georgkordowich/diffpssi 0.0.3, installed by ``__graft_entry__.build()`` into the
git-ignored ``baseline/_ref`` with ``pip install --no-deps --target baseline/_ref <copy of /root/reference>``) on the
bench workload, through the reference's own public API and stock code path:

    PowerSystemSimulation(grid_data=<IEEE 39-bus dict>, parallel_sims=P, solver='heun') -> add_sc_event ->
    per-sim H / D / X_d' set as [P,1] complex attributes (examples/models/ibb_model/ibb_opt.py:48-57) ->
    set_record_function(omega of the 10 generators) -> run()

One process per backend, because the reference fixes its backend in a module-level global at import time
(power_sim_lib/backend.py:8-14): ``--backend numpy`` / ``torch`` (CPU) / ``torch-cuda`` (the reference's own GPU path:
``torch.set_default_device('cuda')``, backend.py:19-21 -- there it is switched by editing DEFAULT_DEVICE; here the same
call is made before the model is built).  None of this repo's kernels, prologue or engine is on this path; only the grid
dict and the seeded parameter draw (data) are shared.

The reported time is the LOOP ONLY, like the reference's own timer (simulator.py:507, 548-550): ``initialize()`` (load
flow, model initialisation, inv(Y) -- O(P^2) in the reference) is timed separately by wrapping it and subtracted.
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, '_ref')


def import_reference(backend):
    """The installed package is ``diffpssi`` but its modules import each other as ``src.diffpssi...`` (the reference is
    meant to be run from its repository root).  A namespace alias ``src`` -> baseline/_ref makes those imports resolve
    without touching the installed files."""
    os.environ['DIFFPSSI_FORCE_SIM_BACKEND'] = 'numpy' if backend == 'numpy' else 'torch'
    os.environ.pop('DIFFPSSI_FORCE_INTEGRATOR', None)
    alias = types.ModuleType('src')
    alias.__path__ = [REF]
    sys.modules['src'] = alias
    import src.diffpssi.power_sim_lib.simulator as simulator
    return simulator


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--backend', default='numpy', choices=('numpy', 'torch', 'torch-cuda'))
    ap.add_argument('--sims', type=int, default=1024)
    ap.add_argument('--fault-steps', type=int, default=0, help='active fault steps inside the sample (0: steps/100)')
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--repeats', type=int, default=1)
    ap.add_argument('--warmup', type=int, default=0)
    ap.add_argument('--threads', type=int, default=0)
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, 'diffpssi')):
        print(json.dumps({'unavailable': 'baseline/_ref is missing (run __graft_entry__.build() where /root/reference exists)'}))
        return
    sys.path.insert(0, ROOT)
    import numpy as np
    # silence the reference's import-time prints (stdout carries the JSON line only)
    real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    simulator = import_reference(args.backend)
    import bench
    from diffpssi_b200 import grids                     # data only: the grid dict in the reference's own schema
    P, dt = args.sims, bench.DT
    cores = os.cpu_count() or 1
    device = 'cpu'
    if args.backend != 'numpy':
        import torch
        if args.threads:
            torch.set_num_threads(args.threads)
        if args.backend == 'torch-cuda':
            if not torch.cuda.is_available():
                real_stdout.write(json.dumps({'unavailable': 'no CUDA device'}) + '\n')
                return
            torch.set_default_device('cuda')            # backend.py:19-21
            device = 'cuda'
        threads = torch.get_num_threads()
    else:
        threads = cores                                 # numpy: whatever its BLAS uses; elementwise ops are single-threaded
    # the sample: ``steps`` consecutive steps with the workload's own fraction of fault steps (20 of 2000 = 1 %); a fault
    # step costs the reference two batched dense inversions (simulator.py:526-528), so the fraction matters
    n_fault = args.fault_steps or max(1, round(args.steps * 20 / bench.T_STEPS))
    k_fault = args.steps // 3
    t0f, t1f = k_fault * dt, (k_fault + n_fault) * dt + 1e-9       # active iff t0f < t <= t1f (static_models.py:46)
    batch = bench.parameter_batch(bench.P_GLOBAL)

    def build():
        sim = simulator.PowerSystemSimulation(time_step=dt, sim_time=args.steps * dt, parallel_sims=P, solver='heun',
                                              grid_data=grids.ieee39(), verbose=False)
        sim.add_sc_event(t0f, t1f, bench.FAULT[2])
        for (g, attr), arr in batch.items():
            col = np.ascontiguousarray(arr[:P]).reshape(P, 1).astype(np.complex128)
            if args.backend != 'numpy':
                col = torch.tensor(col, device=device)
            setattr(sim.get_generator_by_name(g), attr, col)
        sim.set_record_function(lambda s: [s.busses[29 + i].models[0].omega.real for i in range(10)])
        return sim

    sim = build()
    assert len(sim.time) == args.steps
    init_s = [0.0]
    orig_init = sim.initialize

    def timed_init():
        a = time.perf_counter()
        orig_init()
        if device == 'cuda':
            torch.cuda.synchronize()
        init_s[0] = time.perf_counter() - a

    sim.initialize = timed_init
    loops, inits = [], []
    fault_steps = int(sum(1 for t in sim.time if t0f < t <= t1f))
    for i in range(args.warmup + args.repeats):
        if i:
            sim.reset()
        a = time.perf_counter()
        _, res = sim.run()
        if device == 'cuda':
            torch.cuda.synchronize()
        total = time.perf_counter() - a
        if i >= args.warmup:
            loops.append(total - init_s[0])
            inits.append(init_s[0])
    loop = sum(loops) / len(loops)
    res = res.detach().cpu().numpy() if hasattr(res, 'detach') else np.asarray(res)
    out = {'backend': args.backend, 'device': device, 'parallel_sims': P, 'steps': args.steps, 'fault_steps': fault_steps,
           'loop_s': loop, 'loops_s': loops, 'initialize_s': sum(inits) / len(inits), 'sim_steps_per_s': P * args.steps / loop,
           'threads': threads, 'cores': cores, 'finite_fraction': float(np.isfinite(res.real).mean()),
           'omega_absmax': float(np.nanmax(np.abs(res.real)))}
    real_stdout.write(json.dumps(out) + '\n')
    real_stdout.flush()


if __name__ == '__main__':
    main()
