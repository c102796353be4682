#!/usr/bin/env python
"""bench.py -- sim-steps/s of the Pss.run() time loop, IEEE 39-bus / 10 generators (SEXS+TGOV1+STAB1), Heun, fp64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--sims P] [--variant generic]

One bench "step" = one pass of the hot path over the whole synthetic batch: fault re-reduction + the persistent
stepping kernel integrating T = 2000 time steps (10 s at dt = 5 ms, 100 ms three-phase fault at load bus B16) of
``parallel_sims`` = 65 536 trajectories (BASELINE.json configs[2]; sharded by contiguous rows over N GPUs, one process
per GPU, no communication in the loop -- SURVEY.md section 8e).  Metric: sim-steps/s = parallel_sims x T / time of the
loop, initialisation excluded exactly like the reference's own timer (simulator.py:507, 548-550).

JSON line (rank 0): value = device-timed loop with inputs resident in HBM; e2e = the same metric through the public
API (pinned host parameter batch -> H2D -> Pss.run() incl. load flow/initialisation -> recorder D2H into pinned host
memory); roofline = algorithmic fp64 flops of the stepping kernel against max(DFMA, DMMA) peak probes run in the same
process; cpu_baseline = the UNMODIFIED reference (baseline/_ref, installed by __graft_entry__.build()) with its numpy and
torch-CPU backends on the box's host cores on a bounded sample of the same workload, its torch-CUDA path on this GPU as
the incumbent GPU implementation, and the C restatement (oracle/dps_oracle.c, OpenMP) next to them.
``--impl reference`` times the reference itself (the faster of its two CPU backends).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

T_STEPS = 2000
DT = 0.005
FAULT = (1.0, 1.1, 'B16')
P_GLOBAL = 65536
METRIC = 'sim-steps/sec (parallel_sims x steps) IEEE39 Heun fp64'
# algorithmic flops per sim-step,
#   F = S * sum_gens[51 + 2 + 15 pss + 21 avr + 17 gov + 8 g] + U * n_state = 2*10*(106+80) + 6*130
FLOPS_PER_SIM_STEP = 2 * 10 * (51 + 2 + 15 + 21 + 17 + 8 * 10) + 6 * 130
REC_BYTES_PER_SIM_STEP = 8 * 10            # recorder: omega of the 10 generators


def parameter_batch(P_global, seed=0):
    """One global draw (seed 0) of per-sim H, D, X_d' for every generator: H, X_d' ~ U(0.5, 2) x nominal,
    D ~ U(0, 2) -- the reference's identification convention (examples/models/ibb_model/ibb_opt.py:48-57)."""
    from diffpssi_b200 import grids
    rng = np.random.RandomState(seed)
    batch = {}
    for g, (h, xdt) in grids.GENERATOR_NOMINALS['ieee39'].items():
        batch[(g, 'h')] = h * rng.uniform(0.5, 2.0, P_global)
        batch[(g, 'd')] = rng.uniform(0.0, 2.0, P_global)
        batch[(g, 'x_d_t')] = xdt * rng.uniform(0.5, 2.0, P_global)
    return batch


def generic_grid():
    """--variant generic: the same IEEE-39 system with DISTINCT controller settings on every machine (gains and time
    constants spread by a few per cent), so the stepper cannot use its common-constant (constant-bank) specialisation
    nor the two-trajectories-per-thread kernel -- the path a user with a hand-tuned grid gets."""
    from diffpssi_b200 import grids
    grid = grids.ieee39()
    for key, sub, cols in (('avr', 'SEXS', ('K', 'T_a', 'T_b', 'T_e')), ('gov', 'TGOV1', ('R', 'T_1', 'T_3')),
                           ('pss', 'STAB1', ('K', 'T', 'T_1', 'T_3'))):
        table = grid[key][sub]
        head = table[0]
        for i, row in enumerate(table[1:]):
            for c in cols:
                row[head.index(c)] = row[head.index(c)] * (1.0 + 0.01 * (i + 1))
    return grid


def build_sim(P, device, rows, batch_host, variant='headline'):
    """IEEE-39 simulation for rows [rows) of the global parameter draw; parameters handed in as [P,1] complex128
    tensors exactly like a reference user would (ibb_opt.py:48-57)."""
    import torch
    from diffpssi_b200 import Pss, grids
    sim = Pss(parallel_sims=P, sim_time=T_STEPS * DT, time_step=DT, solver='heun',
              grid_data=generic_grid() if variant in ('generic', 'generic-ctrl') else grids.ieee39(), verbose=False, device=device)
    sim.add_sc_event(*FAULT)
    # ('generic-ctrl' / 'generic-rec': only one of the two differences, to see which one costs what)
    if variant in ('generic', 'generic-rec'):          # omega AND delta of every generator: the run-time recorder mask, 160 B per sim-step
        sim.set_record_function(lambda s: [s.busses[29 + i].models[0].omega.real for i in range(10)] +
                                          [s.busses[29 + i].models[0].delta.real for i in range(10)])
    else:
        sim.set_record_function(lambda s: [s.busses[29 + i].models[0].omega.real for i in range(10)])
    assert len(sim.time) == T_STEPS
    set_parameters(sim, batch_host, rows, device)
    return sim


def set_parameters(sim, batch_host, rows, device):
    import torch
    for (g, attr), arr in batch_host.items():
        t = arr[rows[0]:rows[1]] if not torch.is_tensor(arr) else arr
        if not torch.is_tensor(t):
            t = torch.from_numpy(np.ascontiguousarray(t))
        t = t.to(device, non_blocking=True).reshape(-1, 1)
        setattr(sim.get_generator_by_name(g), attr, torch.complex(t, torch.zeros_like(t)))


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def c_port_baseline(sample_sims, batch, threads=None):
    """The C oracle (literal CPU restatement of the reference loop) on the first ``sample_sims`` rows of the workload.
    Loop only; the prologue is excluded like on the GPU side."""
    from oracle import adapter, oracle
    oracle.build()
    threads = threads or os.cpu_count() or 1
    sim = build_sim(sample_sims, 'cpu', (0, sample_sims), batch)
    init = sim.initialize()
    rec = [(0, g, 0) for g in range(10)]
    inp = adapter.oracle_inputs_from_sim(sim, init)
    t0 = time.perf_counter()
    out = oracle.run(inp['n_bus'], inp['gen_bus'], inp['ctrl'], 'heun', DT, inp['fn'], sim.time, inp['events'],
                     inp['raw'], inp['x0'], inp['z0'], rec=rec, full=False, busv=False, threads=threads)
    dt = time.perf_counter() - t0
    return sample_sims * T_STEPS / dt, dt, out['rec']


REF_RUNNER = os.path.join(ROOT, 'baseline', 'run_reference.py')
REF_SIMS = 1024          # the reference's throughput saturates by 1-4 k sims; its initialize() is O(P^2)


def reference_available():
    return os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'diffpssi'))


def run_reference_backend(backend, sims, steps, repeats=1, warmup=0, timeout=900):
    """One process of baseline/run_reference.py (the unmodified reference, loop-only timing) -> its JSON dict.
    The sample covers ``steps`` consecutive steps of the run with the workload's own fault-step fraction (1 %)."""
    cmd = [sys.executable, REF_RUNNER, '--backend', backend, '--sims', str(sims), '--steps', str(steps),
           '--fault-steps', str(max(1, round(steps * 20 / T_STEPS))), '--repeats', str(repeats), '--warmup', str(warmup)]
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {'unavailable': 'timed out after %d s' % timeout}
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith('{'):
            return json.loads(ln)
    return {'unavailable': 'rc %d: %s' % (r.returncode, r.stderr.strip()[-300:])}


def cpu_baseline(batch, with_cuda_incumbent, port_sample):
    """The baselines BASELINE.json north_star names, timed beside the GPU arm on the box's own host cores: the reference's
    numpy and torch-CPU backends (and its torch-CUDA path on this GPU), plus the C port.  ``value`` is the faster of the two
    real CPU backends (kind 'reference'); without baseline/_ref it falls back to the port (kind 'port')."""
    cores = os.cpu_count() or 1
    backends = {}
    v, dt_s, _ = c_port_baseline(port_sample, batch, cores)
    backends['c_port'] = {'sim_steps_per_s': v, 'threads': cores, 'kind': 'port',
                          'sample': '%d trajectories x %d steps, %.1f s of oracle/dps_oracle.c with OpenMP' % (port_sample, T_STEPS, dt_s)}
    best = None
    if reference_available():
        todo = ['numpy', 'torch'] + (['torch-cuda'] if with_cuda_incumbent else [])
        for be in todo:
            d = run_reference_backend(be, REF_SIMS, 300)
            key = {'numpy': 'reference_numpy', 'torch': 'reference_torch_cpu', 'torch-cuda': 'reference_torch_cuda'}[be]
            if 'unavailable' in d:
                backends[key] = d
                continue
            backends[key] = {'sim_steps_per_s': d['sim_steps_per_s'], 'threads': d['threads'], 'kind': 'reference',
                             'device': d['device'], 'loop_s': d['loop_s'], 'initialize_s': d['initialize_s'],
                             'sample': '%d trajectories x %d steps incl. %d fault steps, loop only (simulator.py:507,548-550)'
                                       % (d['parallel_sims'], d['steps'], d['fault_steps'])}
            if be != 'torch-cuda' and (best is None or d['sim_steps_per_s'] > best[1]['sim_steps_per_s']):
                best = (key, d)
    if best is not None:
        key, d = best
        return {'value': d['sim_steps_per_s'], 'unit': 'sim-steps/s', 'cores': cores, 'kind': 'reference',
                'backend': key, 'threads': d['threads'],
                'sample': 'unmodified reference (%s backend), %d trajectories x %d steps of the same workload (rows 0..%d of the '
                          'seed-0 draw), loop only' % (d['backend'], d['parallel_sims'], d['steps'], d['parallel_sims'] - 1),
                'backends': backends}
    return {'value': v, 'unit': 'sim-steps/s', 'cores': cores, 'kind': 'port', 'backend': 'c_port',
            'sample': backends['c_port']['sample'] + ' (baseline/_ref not installed: reference backends not timed)',
            'backends': backends}


def run_reference_arm(args, out=sys.stdout):
    """--impl reference: the reference's own CPU implementation of the path, unmodified (baseline/_ref), through its public
    API: each bench step is one ``run()`` of 1024 trajectories x 100 steps of the workload (1 fault step = the workload's
    1 % fault fraction), loop-only like the reference's own timer.  The faster of its numpy / torch-CPU backends on this
    box is used (probed first); the other one and the C port are reported beside it.  Falls back to the C port (kind
    'port') when baseline/_ref is absent."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    batch = parameter_batch(P_GLOBAL)
    others = {}
    kind, backend, threads = 'port', 'c_port', cores
    if reference_available():
        probe = {be: run_reference_backend(be, REF_SIMS, 40) for be in ('numpy', 'torch')}
        ok = {be: d for be, d in probe.items() if 'unavailable' not in d}
        for be, d in probe.items():
            others['probe_' + be] = d.get('sim_steps_per_s', d.get('unavailable'))
        if ok:
            backend = max(ok, key=lambda be: ok[be]['sim_steps_per_s'])
            d = run_reference_backend(backend, REF_SIMS, 100, repeats=args.steps, warmup=args.warmup, timeout=1500)
            if 'unavailable' not in d:
                kind, threads = 'reference', d['threads']
                sample, ms = REF_SIMS, 1e3 * d['loop_s']
                steps_per_run = d['steps']
                note = ('unmodified reference, %s backend: %d trajectories x %d steps per bench step, loop only; '
                        'initialize() (%.1f s per run, O(P^2)) excluded' % (backend, sample, steps_per_run, d['initialize_s']))
            else:
                others['run_error'] = d['unavailable']
    if kind == 'port':
        sample = int(min(P_GLOBAL, max(cores * 48, 64)))
        times = []
        for i in range(args.warmup + args.steps):
            v, dt, _ = c_port_baseline(sample, batch, cores)
            if i >= args.warmup:
                times.append(dt)
        ms = 1e3 * sum(times) / len(times)
        steps_per_run = T_STEPS
        note = 'CPU restatement of the reference loop (oracle/dps_oracle.c) on a bounded sample'
    else:
        v, dt, _ = c_port_baseline(int(min(P_GLOBAL, max(cores * 48, 64))), batch, cores)
        others['c_port'] = v
    value = sample * steps_per_run / (ms * 1e-3)
    cfg = workload_config(sample, args.gpus, note=note)
    cfg['steps_per_run'] = steps_per_run
    cfg['l2'] = 'n/a (CPU arm)'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'sim-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
        'cpu_baseline': {'value': value, 'unit': 'sim-steps/s', 'cores': cores, 'threads': threads, 'kind': kind,
                         'backend': backend,
                         'sample': '%d trajectories x %d steps per bench step (rows 0..%d of the seed-0 draw)'
                                   % (sample, steps_per_run, sample - 1),
                         'others_sim_steps_per_s': others},
        'e2e': {'value': value, 'unit': 'sim-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    out.write(json.dumps(line) + '\n')
    out.flush()


def kernel_label(info, variant):
    if variant.startswith('generic'):
        consts = 'constant-bank controller constants' if variant == 'generic-rec' else 'per-generator controller constants from shared memory'
        rec = 'omega recorder' if variant == 'generic-ctrl' else 'run-time recorder mask'
        return ('dps_step_pair_kernel<G=10, all controllers, %s, heun, %s> (two trajectories per thread, %d registers)'
                % (consts, rec, info['regs_per_thread']))
    if info['sims_per_block'] % 6 == 0 and info['regs_per_thread'] > 170:
        return 'dps_step_pair_kernel<G=10, heun, omega recorder> (two trajectories per thread, %d registers)' % info['regs_per_thread']
    return ('dps_step_kernel<G=10, all controllers, constant-bank controller constants, heun> (%d registers)'
            % info['regs_per_thread'])


def workload_config(P, n_gpus, note=None, variant='headline', n_rec=10):
    cfg = {'workload': 'IEEE 39-bus New-England, 10 generators each with SEXS+TGOV1+STAB1 (130 states/sim), Heun, '
                       'dt=5 ms, T=2000 steps, 100 ms three-phase fault at bus B16, recorder = omega of 10 generators',
           'parallel_sims': P, 'steps_per_run': T_STEPS, 'solver': 'heun', 'parallelism': 'dp%d (rows of one seed-0 draw)' % n_gpus,
           'parameters': "per-sim H, X_d' ~ U(0.5,2) x nominal and D ~ U(0,2) for every generator",
           'l2': 'recorder output (%.1f GB per bench step per GPU) far exceeds the 126 MB L2, so inputs are evicted '
                 'between timed iterations' % (P / n_gpus * T_STEPS * 8 * n_rec / 1e9)}
    if variant.startswith('generic'):
        cfg['variant'] = ('generic: controller gains / time constants differ per generator (no common-constant kernel '
                          'specialisation, no two-trajectory kernel), recorder = omega and delta of the 10 generators')
    if note:
        cfg['note'] = note
    return cfg


C4_SIMS = 16384
E2E_CHUNKS = int(os.environ.get("DPS_E2E_CHUNKS", 32))      # time slices of the streamed recorder copy (A/B: profiles/round2/README.md)


def run_c4(args, real_stdout):
    """--config c4: BASELINE.json configs[3] -- one parameter-identification step on the IBB model: forward + adjoint
    gradients of the per-trajectory MSE against the reference's golden trajectory with respect to per-sim H, D, X_d' (16 384
    trajectories sharded by rows over N GPUs) plus one parameter shared by all trajectories, whose gradient is all-reduced
    over NCCL; the per-sim loss vector is all-gathered (arg-min of ps_optimization.py:202).  Timed twice: the eager step
    (Python prologue + autograd around the kernels) and the same step replayed from one captured CUDA graph."""
    import torch
    import torch.distributed as dist
    from diffpssi_b200 import Pss, GraphedStep, engine, sharding
    from tests import helpers
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device(dev))
    engine.load_library()
    P_total = args.sims if args.sims != P_GLOBAL else C4_SIMS
    lo, hi = rank * P_total // world, (rank + 1) * P_total // world
    P = hi - lo
    rng = np.random.RandomState(0)
    params = {('Gen 1', 'h'): 3.5 * rng.uniform(0.5, 2, P_total), ('Gen 1', 'd'): rng.uniform(0, 2, P_total),
              ('Gen 1', 'x_d_t'): 0.3 * rng.uniform(0.5, 2, P_total)}
    target = torch.from_numpy(np.load(os.path.join(helpers.GOLD, 'ref_ibb_model.npy'))).to(dev)
    sim = Pss(parallel_sims=P, sim_time=T_STEPS * DT, time_step=DT, solver='heun', grid_data=helpers.load_grid('ibb'),
              verbose=False, device=dev)
    sim.add_sc_event(1, 1.05, 'Bus 1')
    leaves = helpers.apply_overrides(sim, {k: v[lo:hi] for k, v in params.items()}, requires_grad=True)
    shared = torch.tensor([[0.07 + 0j]], dtype=torch.complex128, device=dev, requires_grad=True)
    sim.get_generator_by_name('Gen 1').t_q0_st = shared
    sim.set_record_channels(['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st'])
    every = list(leaves.values()) + [shared]

    def step():
        for l in every:
            l.grad = None
        sim.reset()
        _, res = sim.run()
        loss = ((res - target[None]) ** 2).mean(dim=(1, 2))
        loss.sum().backward()
        return loss.detach()

    def collectives(loss):
        full = sharding.all_gather_rows(loss, P_total)
        sharding.all_reduce_shared_grads([shared])
        return full

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            full = collectives(fn())
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) / n * 1e3
        t = torch.tensor([wall, e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), full

    graphed = GraphedStep(step, warmup=0)       # every execution of step() goes through its private stream (see GraphedStep.eager)
    for _ in range(max(args.warmup, 3)):
        collectives(graphed.eager())
    eager_ms, eager_dev_ms, full_eager = timed(graphed.eager, args.steps)
    grads_eager = [l.grad.clone() for l in every]
    for _ in range(3):
        collectives(graphed())
    graph_ms, graph_dev_ms, full_graph = timed(graphed, args.steps)
    same = bool(torch.equal(torch.nan_to_num(full_eager), torch.nan_to_num(full_graph))) and all(
        bool(torch.equal(torch.nan_to_num(torch.view_as_real(a)), torch.nan_to_num(torch.view_as_real(l.grad))))
        for a, l in zip(grads_eager, every))
    if rank == 0:
        line = {
            'metric': 'sim-steps/sec (parallel_sims x steps) IBB identification step (forward + adjoint) Heun fp64',
            'value': P_total * T_STEPS / (graph_ms * 1e-3), 'unit': 'sim-steps/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': graph_ms, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': "BASELINE configs[3]: IBB, Heun, T=2000, forward + discrete adjoint of the per-sim trajectory "
                                   "MSE w.r.t. per-sim H, D, X_d' and one shared parameter; loss all-gather + NCCL all-reduce of "
                                   "the shared gradient every step", 'parallel_sims': P_total,
                       'parallelism': 'dp%d (rows)' % world, 'step': 'one captured CUDA graph per rank (prologue, '
                       'dps_network_reduce, dps_forward, loss, dps_backward, prologue backward) + 2 collectives'},
            'graphed': {'ms_per_step': graph_ms, 'device_ms_per_step': graph_dev_ms},
            'eager': {'ms_per_step': eager_ms, 'device_ms_per_step': eager_dev_ms,
                      'sim_steps_per_s': P_total * T_STEPS / (eager_ms * 1e-3)},
            'graph_equals_eager': same,
            'best_row': int(torch.argmin(torch.nan_to_num(full_graph, nan=float('inf')))),
            'shared_grad': float(shared.grad.real),
            'gpu_launches': None,
        }
        real_stdout.write(json.dumps(line) + '\n')
        real_stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c5(args, real_stdout):
    """--config c5: BASELINE.json configs[4] -- synthetic tiled 500-bus / 100-generator grid (diffpssi_b200.grids.tiled), all
    three controllers on every machine, RK4, 8192 trajectories x 2000 steps, per-sim H / D / X_d' from the 0.5x..2x range,
    one fault at a load bus, omega of the 100 generators recorded.  The network product V = Z I is 84 % of the flops here
    and runs on the FP64 MMA instruction (dps_step_wide_kernel).  Rows sharded over N GPUs like config 3."""
    import torch
    import torch.distributed as dist
    from diffpssi_b200 import Pss, grids, engine
    from tests import helpers
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device(dev))
    engine.load_library()
    P_total = args.sims if args.sims != P_GLOBAL else 8192
    lo, hi = rank * P_total // world, (rank + 1) * P_total // world
    P, T, n_cells = hi - lo, T_STEPS, 100
    rng = np.random.RandomState(0)
    sim = Pss(parallel_sims=P, sim_time=T * DT, time_step=DT, solver='rk4', grid_data=grids.tiled(n_cells), verbose=False,
              device=dev)
    sim.add_sc_event(1.0, 1.1, 'A3')
    ov = {}
    for c in range(n_cells):
        ov[('M%d' % c, 'h')] = ((6.5 if c % 2 == 0 else 6.175) * rng.uniform(0.5, 2, P_total))[lo:hi]
        ov[('M%d' % c, 'd')] = rng.uniform(0, 2, P_total)[lo:hi]
        ov[('M%d' % c, 'x_d_t')] = (0.3 * rng.uniform(0.5, 2, P_total))[lo:hi]
    helpers.apply_overrides(sim, ov)
    sim.set_record_channels(['M%d.omega' % c for c in range(n_cells)])
    prep = sim.prepare()
    rec = torch.empty((n_cells, T, P), dtype=torch.float64, device=dev)
    xf = torch.empty((n_cells, engine.N_STATES, P), dtype=torch.float64, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        prep.launch(rec, xf)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        prep.launch(rec, xf)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    flops = 4 * n_cells * (51 + 2 + 15 + 21 + 17 + 8 * n_cells) + 14 * 13 * n_cells      # 380 600 per sim-step
    dfma_tf, _ = engine.fp64_peak_probe(8192)
    dmma_tf, _ = engine.fp64_dmma_probe(4096)
    achieved = flops * P * T / (ms * 1e-3) / 1e12
    if rank == 0:
        line = {'metric': 'sim-steps/sec (parallel_sims x steps) tiled 500-bus / 100-gen RK4 fp64', 'value': P_total * T / (ms * 1e-3),
                'unit': 'sim-steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms,
                'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': 'BASELINE configs[4]: tiled 500-bus / 100-generator grid (SEXS+TGOV1+STAB1 on every machine), '
                                       'RK4, dt=5 ms, T=2000, fault at load bus A3, recorder = omega of 100 generators',
                           'parallel_sims': P_total, 'parallelism': 'dp%d (rows)' % world,
                           'l2': 'recorder output (%.1f GB per step per GPU) exceeds the 126 MB L2' % (P * T * n_cells * 8 / 1e9)},
                'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': max(dfma_tf, dmma_tf), 'unit': 'TFLOP/s',
                             'frac': achieved / max(dfma_tf, dmma_tf), 'traffic': None, 'flops_per_sim_step': flops,
                             'kernel': 'dps_step_wide_kernel<all controllers, rk4, FP64 MMA network product>', 'kernel_ms': ms,
                             'peak_dfma_probe': dfma_tf, 'peak_dmma_probe': dmma_tf,
                             'algorithmic_bytes_per_launch': 8 * n_cells * P * T},
                'gpu_launches': 2 * args.steps, 'kernel': prep.plan.info(),
                'finite_fraction': float(torch.isfinite(rec).double().mean()), 'clocks': clocks}
        real_stdout.write(json.dumps(line) + '\n')
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--sims', type=int, default=P_GLOBAL, help='global parallel_sims (default: the BASELINE config)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--variant', default='headline', choices=('headline', 'generic', 'generic-ctrl', 'generic-rec'),
                    help="'generic': distinct controller constants per generator + omega/delta recorder (no specialised kernel)")
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--config', default='c3', choices=('c3', 'c4', 'c5'),
                    help="'c3' (default): the headline IEEE-39 loop; 'c4': the IBB identification step (forward + adjoint + NCCL); "
                         "'c5': the tiled 500-bus / 100-generator RK4 run (FP64-MMA network product)")
    args = ap.parse_args()
    # the JSON line must be the only thing on stdout: libraries (NCCL prints its version with printf) get stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        return run_reference_arm(args, real_stdout)
    if args.config == 'c4':
        return run_c4(args, real_stdout)
    if args.config == 'c5':
        return run_c5(args, real_stdout)

    import torch
    import torch.distributed as dist
    from diffpssi_b200 import engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torchrun --nproc-per-node %d for --gpus %d' % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device(dev))
    engine.load_library()                       # fail loudly if the CUDA engine is missing

    P_total = args.sims
    lo, hi = rank * P_total // world, (rank + 1) * P_total // world
    P = hi - lo
    batch = parameter_batch(P_total)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: value / roofline ----------------------------------------------------------
    sim = build_sim(P, dev, (lo, hi), batch, args.variant)
    prep = sim.prepare()
    n_rec = len(prep.topo['rec'])
    rec = torch.empty((n_rec, T_STEPS, P), dtype=torch.float64, device=dev)
    x_final = torch.empty((prep.topo['n_gen'], engine.N_STATES, P), dtype=torch.float64, device=dev)
    for _ in range(max(args.warmup, 3)):
        prep.launch(rec, x_final)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        prep.reduce_network()
        kev[i][0].record()
        prep.plan.forward(P, prep.x0.detach(), prep.upar.detach(), prep.vpar.detach() if prep.vpar is not None else None,
                          prep.zsub, rec, x_final, None)
        kev[i][1].record()
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = P_total * T_STEPS / (ms_per_step * 1e-3)
    finite = float(torch.isfinite(rec).double().mean())
    # cheap fingerprint of the result (kernel A/B runs must reproduce it): steps 500, 1000, 1999 of every channel
    probe = rec[:, [500, 1000, T_STEPS - 1], :]
    checksum = float(torch.where(torch.isfinite(probe), probe, torch.zeros_like(probe)).abs().sum())
    info = prep.plan.info()

    # fp64 roofline denominator, measured here (MEASURED_PEAKS.json has no fp64 entry)
    dfma_tf, _ = engine.fp64_peak_probe(8192)
    dmma_tf, _ = engine.fp64_dmma_probe(4096)
    peak_tf = max(dfma_tf, dmma_tf)
    achieved_tf = FLOPS_PER_SIM_STEP * P * T_STEPS / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    # DRAM traffic of the dominant kernel, per launch, from the committed ncu capture of this very configuration
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            tj = json.load(f)
        if P_total == P_GLOBAL and world == 1 and info['sims_per_block'] == tj.get('sims_per_block', 12):
            traffic = tj['dram_bytes_per_launch']
    except Exception:
        pass
    rec_bytes = 8 * n_rec                                     # recorder bytes per sim-step (80 for the headline recorder)
    rec_gbs = rec_bytes * P * T_STEPS / (kernel_ms * 1e-3) / 1e9

    # ---------------- end-to-end arm: public API with host buffers ----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        del rec
        torch.cuda.empty_cache()
        from diffpssi_b200 import hostmem
        numa = hostmem.bind_to_gpu_node(local)     # pinned pages on the GPU's own NUMA node (no-op where sysfs does not tell)
        host_params = {k: torch.from_numpy(np.ascontiguousarray(v[lo:hi])).pin_memory() for k, v in batch.items()}
        host_out = torch.empty((n_rec, T_STEPS, P), dtype=torch.float64).pin_memory()
        h2d = sum(v.numel() * 8 for v in host_params.values())
        d2h = host_out.numel() * 8
        n_e2e = 4                   # fixed: ~0.2 s each; independent of --steps so a short run still averages

        def e2e_step():
            set_parameters(sim, host_params, (0, P), dev)           # H2D from pinned host memory
            sim.reset()
            # public API: load flow + initialisation + kernels; the recorder is streamed into pinned host memory
            # ([value][t][batch]) in E2E_CHUNKS time slices while later steps are still being integrated
            _, result = sim.run(host_out=host_out, stream_chunks=E2E_CHUNKS)
            torch.cuda.synchronize()

        for _ in range(3):          # warm-up: first touches of the pinned result buffer and of the allocator's 10 GB block
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {'value': P_total * T_STEPS / float(dt[0]), 'unit': 'sim-steps/s', 'h2d_bytes_per_step': h2d * world,
               'd2h_bytes_per_step': d2h * world, 'ms_per_step': float(dt[0]) * 1e3, 'steps': n_e2e,
               'prologue_ms': sim.timings.get('prologue_s', 0.0) * 1e3, 'stream_chunks': E2E_CHUNKS,
               'numa_node_of_gpu': numa.get('numa_node'),
               'floor': 'device->pinned-host copy rate of this box (profiles/round2/d2h_probe_n8.json): 56.8 GB/s for one rank '
                        'alone, 92.5 GB/s aggregate for 8 ranks at once => 185 ms (N=1) / 113 ms (N=8) for the 10.5 GB of results'}

    # ---------------- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.cuda.synchronize()
        cores = os.cpu_count() or 1
        cpu = cpu_baseline(batch, with_cuda_incumbent=True, port_sample=int(min(P_total, max(cores * 192, 256))))

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'sim-steps/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(P_total, world, variant=args.variant, n_rec=n_rec),
            'roofline': {'bound': 'fp64', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                         'frac': achieved_tf / peak_tf, 'traffic': traffic,
                         'kernel': kernel_label(info, args.variant),
                         'kernel_ms': kernel_ms, 'algorithmic_bytes_per_launch': rec_bytes * P * T_STEPS,
                         'flops_per_sim_step': FLOPS_PER_SIM_STEP,
                         'peak_source': 'max(DFMA probe, DMMA probe) run in this process (same FP64 pipe; nominal 148 SM x 64 lanes x '
                                        '2 x 1.965 GHz = 37.2); MEASURED_PEAKS.json has no fp64 entry and the profiling guide '
                                        'states no fp64 fallback',
                         'peak_dfma_probe': dfma_tf, 'peak_dmma_probe': dmma_tf,
                         'hbm': {'achieved': rec_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': rec_gbs / hbm_peak,
                                 'bytes_per_sim_step': rec_bytes}},
            'cpu_baseline': cpu, 'e2e': e2e,
            'gpu_launches': 2 * args.steps,
            'kernel': info, 'finite_fraction': finite, 'checksum': checksum, 'clocks': clocks,
        }
        real_stdout.write(json.dumps(line) + '\n')
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
