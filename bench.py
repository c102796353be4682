#!/usr/bin/env python
"""bench.py -- sim-steps/s of the Pss.run() time loop, IEEE 39-bus / 10 generators (SEXS+TGOV1+STAB1), Heun, fp64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--sims P]

One bench "step" = one pass of the hot path over the whole synthetic batch: fault re-reduction + the persistent
stepping kernel integrating T = 2000 time steps (10 s at dt = 5 ms, 100 ms three-phase fault at load bus B16) of
``parallel_sims`` = 65 536 trajectories (BASELINE.json configs[2]; sharded by contiguous rows over N GPUs, one process
per GPU, no communication in the loop -- SURVEY.md section 8e).  Metric: sim-steps/s = parallel_sims x T / time of the
loop, initialisation excluded exactly like the reference's own timer (simulator.py:507, 548-550).

JSON line (rank 0): value = device-timed loop with inputs resident in HBM; e2e = the same metric through the public
API (pinned host parameter batch -> H2D -> Pss.run() incl. load flow/initialisation -> recorder D2H into pinned host
memory); roofline = algorithmic fp64 flops of the stepping kernel against a DFMA peak measured in the same process;
cpu_baseline = the C oracle (oracle/dps_oracle.c, OpenMP over all host cores) on a bounded sample of the same workload.
``--impl reference`` times that CPU restatement alone (the reference is pure Python and does not travel to the GPU box;
DESIGN.md section 7).
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


def build_sim(P, device, rows, batch_host):
    """IEEE-39 simulation for rows [rows) of the global parameter draw; parameters handed in as [P,1] complex128
    tensors exactly like a reference user would (ibb_opt.py:48-57)."""
    import torch
    from diffpssi_b200 import Pss, grids
    sim = Pss(parallel_sims=P, sim_time=T_STEPS * DT, time_step=DT, solver='heun', grid_data=grids.ieee39(),
              verbose=False, device=device)
    sim.add_sc_event(*FAULT)
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


def cpu_baseline(sample_sims, batch, threads=None):
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


def run_reference_arm(args, out=sys.stdout):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure Python/numpy/torch and
    is not present on the GPU box, so this arm times its C restatement (oracle, kind 'port') with every host thread."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    batch = parameter_batch(P_GLOBAL)
    # size the per-step sample for ~2 s of CPU work: ~40 ms per trajectory per core (measured), bounded
    sample = int(min(P_GLOBAL, max(cores * 48, 64)))
    times = []
    for i in range(args.warmup + args.steps):
        v, dt, _ = cpu_baseline(sample, batch, cores)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = sample * T_STEPS / (ms * 1e-3)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'sim-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(sample, args.gpus, note='CPU restatement of the reference loop on a bounded sample'),
        'cpu_baseline': {'value': value, 'unit': 'sim-steps/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d trajectories x %d steps per bench step (rows 0..%d of the seed-0 draw)'
                                   % (sample, T_STEPS, sample - 1)},
        'e2e': {'value': value, 'unit': 'sim-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    out.write(json.dumps(line) + '\n')
    out.flush()


def workload_config(P, n_gpus, note=None):
    cfg = {'workload': 'IEEE 39-bus New-England, 10 generators each with SEXS+TGOV1+STAB1 (130 states/sim), Heun, '
                       'dt=5 ms, T=2000 steps, 100 ms three-phase fault at bus B16, recorder = omega of 10 generators',
           'parallel_sims': P, 'steps_per_run': T_STEPS, 'solver': 'heun', 'parallelism': 'dp%d (rows of one seed-0 draw)' % n_gpus,
           'parameters': "per-sim H, X_d' ~ U(0.5,2) x nominal and D ~ U(0,2) for every generator",
           'l2': 'recorder output (%.1f GB per bench step per GPU) far exceeds the 126 MB L2, so inputs are evicted '
                 'between timed iterations' % (P / n_gpus * T_STEPS * REC_BYTES_PER_SIM_STEP / 1e9)}
    if note:
        cfg['note'] = note
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--sims', type=int, default=P_GLOBAL, help='global parallel_sims (default: the BASELINE config)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    # the JSON line must be the only thing on stdout: libraries (NCCL prints its version with printf) get stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        return run_reference_arm(args, real_stdout)

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
    sim = build_sim(P, dev, (lo, hi), batch)
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
    peak_tf, _ = engine.fp64_peak_probe(8192)
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
    rec_gbs = REC_BYTES_PER_SIM_STEP * P * T_STEPS / (kernel_ms * 1e-3) / 1e9

    # ---------------- end-to-end arm: public API with host buffers ----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        del rec
        torch.cuda.empty_cache()
        host_params = {k: torch.from_numpy(np.ascontiguousarray(v[lo:hi])).pin_memory() for k, v in batch.items()}
        host_out = torch.empty((n_rec, T_STEPS, P), dtype=torch.float64).pin_memory()
        h2d = sum(v.numel() * 8 for v in host_params.values())
        d2h = host_out.numel() * 8
        n_e2e = 4                   # fixed: ~0.2 s each; independent of --steps so a short run still averages

        def e2e_step():
            set_parameters(sim, host_params, (0, P), dev)           # H2D from pinned host memory
            sim.reset()
            # public API: load flow + initialisation + kernels; the recorder is streamed into pinned host memory
            # ([value][t][batch]) in 8 time slices while later steps are still being integrated
            _, result = sim.run(host_out=host_out, stream_chunks=8)
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
               'prologue_ms': sim.timings.get('prologue_s', 0.0) * 1e3}

    # ---------------- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = int(min(P_total, max(cores * 192, 256)))
        v, dt_s, cpu_rec = cpu_baseline(sample, batch, cores)
        cpu = {'value': v, 'unit': 'sim-steps/s', 'cores': cores, 'kind': 'port',
               'sample': '%d trajectories x %d steps (rows 0..%d of the same seed-0 draw), %.1f s of C oracle with '
                         'OpenMP on %d threads' % (sample, T_STEPS, sample - 1, dt_s, cores)}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'sim-steps/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(P_total, world),
            'roofline': {'bound': 'fp64', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                         'frac': achieved_tf / peak_tf, 'traffic': traffic,
                         'kernel': ('dps_step_pair_kernel<G=10, heun, omega recorder> (two trajectories per thread)'
                                    if info['sims_per_block'] == 24 else
                                    'dps_step_kernel<G=10, all controllers, constant-bank controller constants, heun>'),
                         'kernel_ms': kernel_ms, 'algorithmic_bytes_per_launch': REC_BYTES_PER_SIM_STEP * P * T_STEPS,
                         'flops_per_sim_step': FLOPS_PER_SIM_STEP,
                         'peak_source': 'DFMA probe (dps_fp64_peak_probe) in this process; MEASURED_PEAKS.json has no fp64 entry',
                         'hbm': {'achieved': rec_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': rec_gbs / hbm_peak,
                                 'bytes_per_sim_step': REC_BYTES_PER_SIM_STEP}},
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
