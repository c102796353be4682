"""TEST INFRASTRUCTURE -- generates the fixtures under tests/golden/ (run in the build container only).

    python oracle/gen_golden.py            # everything (a few minutes on 8 cores)
    python oracle/gen_golden.py k2a        # cases whose name contains 'k2a'

It imports the unmodified reference from /root/reference through ``oracle/ref_harness.py`` with the
reference's *torch* backend (the path parity is judged against, BASELINE.json north_star) and stores

  tests/golden/grids/<grid>.json      the example grid dicts (reference: examples/models/*/*_model.py ``load()``)
  tests/golden/ref_<name>.npy         copies of the reference's own golden arrays
                                      (reference: tests/integration/data/*.npy, made by its numpy backend)
  tests/golden/<case>.npz             live torch-backend outputs: snapshot after initialize(), every state /
                                      algebraic / bus voltage at a subset of steps (all steps around the fault),
                                      the example's own record channels at every step, and for ``grad_*`` cases
                                      the autograd gradients of a per-sim MSE loss.

``/root/reference`` does not exist on the GPU box; tests read only the committed fixtures.
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

GRIDS = {
    'ibb': ('ibb_model', 'ibb_model'),
    'ibb_trans': ('ibb_transformer', 'ibb_trans_model'),
    'ibb_wc': ('ibb_with_controllers', 'ibb_wc_model'),
    'ieee9': ('ieee_9bus', 'ieee_9bus_model'),
    'k2a': ('k2a', 'k2a_model'),
}

# (grid, sim_time, events, reference golden file, record channels of the example's *_sim.py)
EXAMPLES = {
    'ibb': (10, [(1, 1.05, 'Bus 1')], 'ibb_model',
            ['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st']),                     # ibb_sim.py:20-24,41
    'ibb_trans': (10, [(1, 1.05, 'Bus 0')], 'ibb_transformer',
                  ['G2.omega', 'G2.e_q_st', 'G2.e_d_st']),                        # ibb_trans_sim.py:20-24,41
    'ibb_wc': (10, [(1, 1.05, 'B2')], 'ibb_with_controllers',
               ['G1.omega']),                                                     # golden holds omega only
    'ieee9': (5, [(1, 1.05, 'B8')], 'ieee_9bus',
              ['G1.omega', 'G1.p_e', 'G1.v_re', 'G1.v_im',
               'G2.omega', 'G2.p_e', 'bus:B2.v_re', 'bus:B2.v_im',
               'G3.omega', 'G3.p_e', 'bus:B3.v_re', 'bus:B3.v_im']),              # ieee_9bus_sim.py:20-36,52
    'k2a': (10, [(1, 1.1, 'B1')], 'k2a',
            ['G%d.%s' % (g, s) for s in ('delta', 'omega', 'e_q_st', 'e_d_st') for g in (1, 2, 3, 4)]),
}

DT = 0.005


def keep_steps(T, events, stride=5, pad=12):
    """Steps stored for the all-channel trajectory: every ``stride``-th, the first 3, the last, and every step
    from ``pad`` before a fault to ``pad`` after it."""
    keep = set(range(0, T, stride)) | {0, 1, 2, T - 1}
    for (t0, t1, _b) in events:
        a = int(round(t0 / DT)) - pad
        b = int(round(t1 / DT)) + pad
        keep |= set(range(max(a, 0), min(b, T)))
    return np.array(sorted(keep), dtype=np.int64)


def save_case(name, t, rec, names, snap, events, solver, extra=None, stride=5):
    T = rec.shape[1]
    steps = keep_steps(T, events, stride=stride)
    out = {
        'steps': steps,
        'traj': rec[:, steps, :],                  # [P, n_keep, C]
        'names': np.array(names),
        'time': t,
        'solver': np.array(solver),
        'events': np.array(json.dumps(events)),
    }
    for k, v in snap.items():
        out['init/' + k] = v
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)
    print('wrote', name, 'P=%d T=%d C=%d kept=%d' % (rec.shape[0], T, rec.shape[2], len(steps)))


def example_cases(sel):
    for gname, (sim_time, events, ref_file, chans) in EXAMPLES.items():
        grid = rh.load_example_grid(*GRIDS[gname])
        for solver in ('heun', 'euler', 'rk4'):
            name = '%s_%s' % (gname, solver)
            if sel and not any(s in name for s in sel):
                continue
            sim, _ = rh.build(grid, 1, solver, DT, sim_time, events)
            t, rec, names, snap, _ = rh.run_full(sim)
            idx = [names.index(c) for c in chans]
            extra = {'example_channels': np.array(chans), 'example_rec': rec[0][:, idx]}
            if solver == 'heun':
                ref = np.load(os.path.join(rh.REFERENCE_ROOT, 'tests', 'integration', 'data', ref_file + '.npy'))
                d = np.abs(ref - rec[0][:, idx]).max()
                print('   torch-vs-reference-golden(%s) max|diff| = %.3e' % (ref_file, d))
                assert d < 1e-6      # the reference's own tolerance, tests/integration/test_example_sims.py:18-19
            save_case(name, t, rec, names, snap, events, solver, extra)


def batch_cases(sel):
    rng = np.random.RandomState(0)
    # --- K2A, per-sim H / D / X_d' on every generator, narrow (stable) range, plus PSS gains on G1
    name = 'k2a_batch_heun'
    if not sel or any(s in name for s in sel):
        grid = rh.load_example_grid(*GRIDS['k2a'])
        P = 6
        nominal = {'h': [6.5, 6.5, 6.175, 6.175], 'x_d_t': [0.3] * 4}
        ov = {}
        for gi in range(4):
            gn = 'G%d' % (gi + 1)
            ov[('gen', gn, 'h')] = nominal['h'][gi] * rng.uniform(0.9, 1.1, P)
            ov[('gen', gn, 'd')] = rng.uniform(0.0, 2.0, P)
            ov[('gen', gn, 'x_d_t')] = nominal['x_d_t'][gi] * rng.uniform(0.9, 1.1, P)
        ov[('gen', 'G1', 'stabilizer.washout.k_w')] = 50.0 * rng.uniform(0.5, 2.0, P)
        ov[('gen', 'G2', 'exciter.pt1_lim.gain_pt1')] = 100.0 * rng.uniform(0.8, 1.2, P)
        events = [(1, 1.1, 'B1')]
        sim, _ = rh.build(grid, P, 'heun', DT, 4, events, ov)
        t, rec, names, snap, _ = rh.run_full(sim)
        extra = {'override/' + '|'.join(k): v for k, v in ov.items()}
        save_case(name, t, rec, names, snap, events, 'heun', extra)
    # --- IBB, the reference's own 0.5x..2x identification range (ibb_opt.py:48-57); contains unstable rows
    name = 'ibb_batch_heun'
    if not sel or any(s in name for s in sel):
        grid = rh.load_example_grid(*GRIDS['ibb'])
        P = 8
        ov = {('gen', 'Gen 1', 'h'): 3.5 * rng.uniform(0.5, 2.0, P),
              ('gen', 'Gen 1', 'd'): rng.uniform(0.0, 2.0, P),
              ('gen', 'Gen 1', 'x_d_t'): 0.3 * rng.uniform(0.5, 2.0, P)}
        events = [(1, 1.05, 'Bus 1')]
        sim, _ = rh.build(grid, P, 'heun', DT, 10, events, ov)
        t, rec, names, snap, _ = rh.run_full(sim)
        extra = {'override/' + '|'.join(k): v for k, v in ov.items()}
        save_case(name, t, rec, names, snap, events, 'heun', extra)
    # --- per-sim network: x_d_st (enters Y, synchronous_machine.py:336) varied -> per-sim Z path
    name = 'ieee9_persimz_heun'
    if not sel or any(s in name for s in sel):
        grid = rh.load_example_grid(*GRIDS['ieee9'])
        P = 4
        ov = {('gen', 'G2', 'x_d_st'): 0.2 * rng.uniform(0.8, 1.2, P),
              ('gen', 'G3', 'h'): 2.766544 * rng.uniform(0.8, 1.2, P)}
        events = [(1, 1.05, 'B8')]
        sim, _ = rh.build(grid, P, 'heun', DT, 3, events, ov)
        t, rec, names, snap, _ = rh.run_full(sim)
        extra = {'override/' + '|'.join(k): v for k, v in ov.items()}
        save_case(name, t, rec, names, snap, events, 'heun', extra)
    # --- two overlapping / cancelling events: pins the `else` quirk of simulator.py:524-531
    name = 'ieee9_two_events_heun'
    if not sel or any(s in name for s in sel):
        grid = rh.load_example_grid(*GRIDS['ieee9'])
        events = [(0.5, 0.6, 'B8'), (0.55, 0.7, 'B5')]
        sim, _ = rh.build(grid, 1, 'heun', DT, 2, events)
        t, rec, names, snap, _ = rh.run_full(sim)
        save_case(name, t, rec, names, snap, events, 'heun', stride=1)
    name = 'k2a_two_events_rk4'
    if not sel or any(s in name for s in sel):
        grid = rh.load_example_grid(*GRIDS['k2a'])
        events = [(0.2, 0.3, 'B7'), (0.1, 0.25, 'B1')]
        sim, _ = rh.build(grid, 1, 'rk4', DT, 1.5, events)
        t, rec, names, snap, _ = rh.run_full(sim)
        save_case(name, t, rec, names, snap, events, 'rk4', stride=1)


def param_event_cases(sel):
    """ParamEvent (static_models.py:52-95): an attribute is overwritten mid-run, nothing is re-initialised."""
    name = 'ibb_param_event_heun'
    if not sel or any(s in name for s in sel):
        grid = rh.load_example_grid(*GRIDS['ibb'])
        events = [(0.5, 0.55, 'Bus 1')]
        pev = [(1.0, ('gen', 'Gen 1', 'd'), 2.0), (1.5, ('gen', 'Gen 1', 'h'), 5.0)]
        sim, _ = rh.build(grid, 2, 'heun', DT, 3.0, events, param_events=pev)
        t, rec, names, snap, _ = rh.run_full(sim)
        extra = {'param_events': np.array(json.dumps([[t_, '|'.join(p_[1:]), v_] for t_, p_, v_ in pev]))}
        save_case(name, t, rec, names, snap, events, 'heun', extra, stride=2)
    name = 'k2a_param_event_rk4'
    if not sel or any(s in name for s in sel):
        grid = rh.load_example_grid(*GRIDS['k2a'])
        events = [(0.2, 0.3, 'B1')]
        pev = [(0.5, ('gen', 'G1', 'exciter.pt1_lim.gain_pt1'), 50.0), (0.5, ('gen', 'G3', 'stabilizer.washout.k_w'), 20.0),
               (1.0, ('gen', 'G2', 'governor.d_t'), 0.5)]
        sim, _ = rh.build(grid, 1, 'rk4', DT, 2.0, events, param_events=pev)
        t, rec, names, snap, _ = rh.run_full(sim)
        extra = {'param_events': np.array(json.dumps([[t_, '|'.join(p_[1:]), v_] for t_, p_, v_ in pev]))}
        save_case(name, t, rec, names, snap, events, 'rk4', extra, stride=2)


def grad_cases(sel):
    """Gradient known-answer vectors.  The reference's tests pin no gradient values
    (tests/integration/test_example_opts.py asserts nothing), so these come from the reference's own
    autograd (``loss.sum().backward()``, ps_optimization.py:212) and are cross-checked here by central
    finite differences on the reference itself."""
    import torch
    rng = np.random.RandomState(1)

    def one(name, gname, P, solver, sim_time, events, ov, chans, fd_check, pev=None):
        if sel and not any(s in name for s in sel):
            return
        grid = rh.load_example_grid(*GRIDS[gname])
        ref_traj = np.load(os.path.join(GOLD, '%s_%s.npz' % (gname, 'heun')))
        all_names = list(ref_traj['names'])

        def run(ovv, grad):
            sim, leaves = rh.build(grid, P, solver, DT, sim_time, events, ovv, requires_grad=grad, param_events=pev)
            t, rec, names, snap, rec_t = rh.run_full(sim)
            idx = [names.index(c) for c in chans]
            return sim, leaves, rec_t[:, :, idx], names, snap, t, rec

        sim, leaves, out, names, snap, t, rec = run(ov, True)
        T = out.shape[1]
        # target: a smooth synthetic signal (deterministic) so the loss is O(1e-3) and well conditioned
        tt = torch.tensor(t)[None, :, None]
        target = 0.002 * torch.sin(2.0 * tt) * torch.ones_like(out.real)
        loss = ((out.real - target) ** 2).mean(dim=(1, 2))          # [P]
        loss.sum().backward()
        grads = {k: v.grad.real.numpy().reshape(-1).copy() for k, v in leaves.items()}
        extra = {'override/' + '|'.join(k): v for k, v in ov.items()}
        extra.update({'grad/' + '|'.join(k): v for k, v in grads.items()})
        extra['loss'] = loss.detach().numpy()
        extra['loss_channels'] = np.array(chans)
        extra['target'] = target[0, :, 0].numpy()
        if pev:
            extra['param_events'] = np.array(json.dumps([[t_, '|'.join(p_[1:]), v_] for t_, p_, v_ in pev]))
        # finite-difference cross-check of the oracle's own gradients (one parameter, all sims at once: rows
        # are independent so a simultaneous perturbation of every row gives every row's derivative)
        if fd_check:
            k0 = list(ov.keys())[0]
            h = 1e-6 * np.abs(ov[k0]).max()
            lp = []
            for sgn in (+1, -1):
                ov2 = dict(ov)
                ov2[k0] = ov[k0] + sgn * h
                _, _, o2, _, _, _, _ = run(ov2, False)
                lp.append(((o2.real - target) ** 2).mean(dim=(1, 2)).numpy())
            fd = (lp[0] - lp[1]) / (2 * h)
            rel = np.abs(fd - grads[k0]) / (np.abs(grads[k0]) + 1e-30)
            print('   FD check on %s: max rel diff %.2e' % ('|'.join(k0), rel.max()))
            extra['fd/' + '|'.join(k0)] = fd
        save_case(name, t, rec, names, snap, events, solver, extra, stride=20)

    P = 4
    one('grad_ibb_heun', 'ibb', P, 'heun', 2.0, [(1, 1.05, 'Bus 1')],
        {('gen', 'Gen 1', 'h'): 3.5 * rng.uniform(0.8, 1.5, P),
         ('gen', 'Gen 1', 'd'): rng.uniform(0.5, 2.0, P),
         ('gen', 'Gen 1', 'x_d_t'): 0.3 * rng.uniform(0.8, 1.25, P),
         ('gen', 'Gen 1', 'x_q'): 1.76 * rng.uniform(0.8, 1.25, P)},
        ['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st'], True)
    one('grad_ibb_euler', 'ibb', P, 'euler', 1.5, [(1, 1.05, 'Bus 1')],
        {('gen', 'Gen 1', 'h'): 3.5 * rng.uniform(0.8, 1.5, P),
         ('gen', 'Gen 1', 'x_d'): 1.81 * rng.uniform(0.8, 1.25, P)},
        ['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st'], False)
    one('grad_ibb_rk4', 'ibb', P, 'rk4', 1.5, [(1, 1.05, 'Bus 1')],
        {('gen', 'Gen 1', 'h'): 3.5 * rng.uniform(0.8, 1.5, P),
         ('gen', 'Gen 1', 't_d0_t'): 8.0 * rng.uniform(0.8, 1.25, P)},
        ['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st'], False)
    # gradients THROUGH parameter events (static_models.py:52-95): D is overwritten at t = 1.0 (its leaf only sees the first
    # second), H at t = 1.5; X_d' keeps its leaf for the whole run
    rng_pe = np.random.RandomState(11)           # (its own stream: the cases below keep their draws)
    one('grad_ibb_param_event_heun', 'ibb', P, 'heun', 2.5, [(0.5, 0.55, 'Bus 1')],
        {('gen', 'Gen 1', 'h'): 3.5 * rng_pe.uniform(0.8, 1.5, P),
         ('gen', 'Gen 1', 'd'): rng_pe.uniform(0.5, 2.0, P),
         ('gen', 'Gen 1', 'x_d_t'): 0.3 * rng_pe.uniform(0.8, 1.25, P)},
        ['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st'], True,
        pev=[(1.0, ('gen', 'Gen 1', 'd'), 2.0), (1.5, ('gen', 'Gen 1', 'h'), 5.0)])
    P = 3
    # H_lim is raised to 0.3 so the PSS works in its linear range part of the time (with the stock 0.03 it is
    # clamped for the whole post-fault window and every PSS gradient is exactly zero).
    one('grad_k2a_heun', 'k2a', P, 'heun', 2.0, [(0.1, 0.13, 'B1')],
        {('gen', 'G1', 'stabilizer.washout.k_w'): 50.0 * rng.uniform(0.5, 2.0, P),
         ('gen', 'G1', 'stabilizer.lead_lag1.t_1'): 0.5 * rng.uniform(0.5, 2.0, P),
         ('gen', 'G1', 'stabilizer.limiter.limit'): 0.3 * np.ones(P),
         ('gen', 'G2', 'h'): 6.5 * rng.uniform(0.8, 1.25, P),
         ('gen', 'G3', 'exciter.pt1_lim.gain_pt1'): 100.0 * rng.uniform(0.8, 1.25, P),
         ('gen', 'G4', 'governor.lead_lag.t_1'): 1.0 * rng.uniform(0.8, 1.25, P)},
        ['G%d.omega' % g for g in (1, 2, 3, 4)] + ['G1.e_q_st', 'G3.e_d_st'], True)
    network_grad_cases(sel, one)


def network_grad_cases(sel, one):
    """Leaves that enter the dynamic admittance matrix (x_d'' of a generator: synchronous_machine.py:336): the gradient
    flows through Z = inv(Y_dyn) and through the fault re-reduction (simulator.py:386, 526-528).

    The unmodified reference cannot differentiate this combination: ``Y = inv(Z); Y[:, b, b] = 1e6`` writes in place
    into the output of ``inv``, which its backward needs ("modified by an inplace operation").  For these cases only,
    ``torch.linalg.inv`` is wrapped to return a clone -- an identity for the values and for the mathematics of the
    gradient -- and the result is cross-checked by central finite differences of the forward pass, whose values the
    wrapper cannot change."""
    import torch
    rng = np.random.RandomState(7)
    orig_inv = torch.linalg.inv
    torch.linalg.inv = lambda a, **kw: orig_inv(a, **kw).clone()
    try:
        _network_grad_cases(sel, one, rng)
    finally:
        torch.linalg.inv = orig_inv


def _network_grad_cases(sel, one, rng):
    P = 4
    one('grad_ibb_z_heun', 'ibb', P, 'heun', 2.0, [(1, 1.05, 'Bus 1')],
        {('gen', 'Gen 1', 'x_d_st'): 0.23 * rng.uniform(0.8, 1.25, P),
         ('gen', 'Gen 1', 'h'): 3.5 * rng.uniform(0.8, 1.5, P)},
        ['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st'], True)
    P = 3
    one('grad_k2a_z_heun', 'k2a', P, 'heun', 1.0, [(0.1, 0.13, 'B1')],
        {('gen', 'G1', 'x_d_st'): 0.25 * rng.uniform(0.8, 1.25, P),
         ('gen', 'G3', 'x_d_st'): 0.25 * rng.uniform(0.8, 1.25, P),
         ('gen', 'G2', 'h'): 6.5 * rng.uniform(0.8, 1.25, P)},
        ['G%d.omega' % g for g in (1, 2, 3, 4)] + ['G1.e_q_st', 'G3.e_d_st'], True)


def headline_cases(sel):
    """Live reference outputs on the BASELINE configurations the bench times (task 1): the builder-authored
    IEEE 39-bus and tiled grids (diffpssi_b200/grids.py, written in the reference's dict schema so that the unmodified
    reference loads them), at full run length."""
    from diffpssi_b200 import grids
    import bench

    def dump_grid(name, grid):
        with open(os.path.join(GOLD, 'grids', name + '.json'), 'w') as f:
            json.dump(grid, f)

    # --- config 3 grid, nominal parameters, full 2000 Heun steps, the bench's fault
    name = 'ieee39_heun'
    events = [tuple(bench.FAULT)]
    if not sel or any(s in name for s in sel):
        dump_grid('ieee39', grids.ieee39())
        sim, _ = rh.build(grids.ieee39(), 1, 'heun', DT, bench.T_STEPS * DT, events)
        t, rec, names, snap, _ = rh.run_full(sim)
        save_case(name, t, rec, names, snap, events, 'heun', stride=10)
    # --- the first rows of the bench's own seed-0 parameter draw (H, X_d' ~ U(0.5,2) x nominal, D ~ U(0,2))
    name = 'ieee39_batch_heun'
    if not sel or any(s in name for s in sel):
        dump_grid('ieee39', grids.ieee39())
        P = 4
        batch = bench.parameter_batch(bench.P_GLOBAL)
        ov = {('gen', g, attr): arr[:P].copy() for (g, attr), arr in batch.items()}
        sim, _ = rh.build(grids.ieee39(), P, 'heun', DT, bench.T_STEPS * DT, events, ov)
        t, rec, names, snap, _ = rh.run_full(sim)
        extra = {'override/' + '|'.join(k): v for k, v in ov.items()}
        save_case(name, t, rec, names, snap, events, 'heun', extra, stride=10)
    # --- config 5 family: tiled grid, RK4, per-sim H / X_d', fault at a load bus
    for name, n_cells, P, T in (('tiled20_rk4', 20, 2, 400), ('tiled100_rk4', 100, 1, 300)):
        if sel and not any(s in name for s in sel):
            continue
        grid = grids.tiled(n_cells)
        dump_grid('tiled%d' % n_cells, grid)
        rng = np.random.RandomState(50 + n_cells)
        ov = {}
        for c in range(n_cells):
            ov[('gen', 'M%d' % c, 'h')] = (6.5 if c % 2 == 0 else 6.175) * rng.uniform(0.8, 1.25, P)
            ov[('gen', 'M%d' % c, 'x_d_t')] = 0.3 * rng.uniform(0.9, 1.1, P)
        events5 = [(0.25, 0.35, 'A3')]
        sim, _ = rh.build(grid, P, 'rk4', DT, T * DT, events5, ov)
        t, rec, names, snap, _ = rh.run_full(sim)
        # generator states/algebraics of every machine, bus voltages of the first two cells and the faulted neighbourhood
        keep = [i for i, n in enumerate(names) if not n.startswith('bus:') or n.split('.')[0][4:] in
                ('G0', 'H0', 'A0', 'B0', 'X0', 'A3', 'H3', 'X3', 'X%d' % (n_cells - 1))]
        extra = {'override/' + '|'.join(k): v for k, v in ov.items()}
        save_case(name, t, rec[:, :, keep], [names[i] for i in keep], snap, events5, 'rk4', extra, stride=10)


def ident_range_grad_case(sel):
    """BASELINE config 4 as the reference itself would run it (examples/models/ibb_model/ibb_opt.py:48-57,
    optimization_lib/ps_optimization.py:176-212): IBB, Heun, the full 10 s, per-sim H / D / X_d' leaves drawn from the
    0.5x..2x identification range, loss = per-sim MSE of (omega, e_q'', e_d'') against the reference's own golden
    trajectory (tests/integration/data/ibb_model.npy), gradients from ``loss.sum().backward()`` (``p.grad.real``)."""
    import torch
    name = 'grad_ibb_range_heun'
    if sel and not any(s in name for s in sel):
        return
    rng = np.random.RandomState(4)
    P = 8
    ov = {('gen', 'Gen 1', 'h'): 3.5 * rng.uniform(0.5, 2.0, P),
          ('gen', 'Gen 1', 'd'): rng.uniform(0.0, 2.0, P),
          ('gen', 'Gen 1', 'x_d_t'): 0.3 * rng.uniform(0.5, 2.0, P)}
    grid = rh.load_example_grid(*GRIDS['ibb'])
    events = [(1, 1.05, 'Bus 1')]
    chans = ['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st']
    sim, leaves = rh.build(grid, P, 'heun', DT, 10, events, ov, requires_grad=True)
    t, rec, names, snap, rec_t = rh.run_full(sim)
    idx = [names.index(c) for c in chans]
    target = torch.tensor(np.load(os.path.join(rh.REFERENCE_ROOT, 'tests', 'integration', 'data', 'ibb_model.npy')))
    out = rec_t[:, :, idx]
    loss = ((out.real - target[None]) ** 2).mean(dim=(1, 2))
    loss.sum().backward()
    extra = {'override/' + '|'.join(k): v for k, v in ov.items()}
    extra.update({'grad/' + '|'.join(k): v.grad.real.numpy().reshape(-1).copy() for k, v in leaves.items()})
    extra['loss'] = loss.detach().numpy()
    extra['loss_channels'] = np.array(chans)
    extra['target2d'] = target.numpy()
    # finite-difference cross-check of every leaf (rows are independent: perturb all rows at once)
    for k0 in ov:
        h = 1e-6 * np.abs(ov[k0]).max()
        lp = []
        for sgn in (+1, -1):
            ov2 = dict(ov)
            ov2[k0] = ov[k0] + sgn * h
            s2, _ = rh.build(grid, P, 'heun', DT, 10, events, ov2)
            _, _, n2, _, r2 = rh.run_full(s2)
            lp.append(((r2[:, :, idx].real - target[None]) ** 2).mean(dim=(1, 2)).numpy())
        fd = (lp[0] - lp[1]) / (2 * h)
        g = extra['grad/' + '|'.join(k0)]
        print('   FD check on %s: max rel diff %.2e (scale %.2e)' % ('|'.join(k0), np.abs(fd - g).max() / np.abs(g).max(), np.abs(g).max()))
        extra['fd/' + '|'.join(k0)] = fd
    save_case(name, t, rec, names, snap, events, 'heun', extra, stride=20)


def _stub_matplotlib():
    """optimization_lib/ps_optimization.py:7 imports matplotlib, which this image does not have; its plotting is only
    reached with enable_plots=True / at the end of run().  A do-nothing stand-in lets the unmodified module import."""
    import types
    if 'matplotlib' in sys.modules:
        return
    mpl = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')

    class _Anything(object):
        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

        def __iter__(self):
            return iter(())

    def _plt_attr(name):
        if name.startswith('__'):
            raise AttributeError(name)
        return lambda *a, **k: _Anything()
    plt.__getattr__ = _plt_attr
    mpl.pyplot = plt
    sys.modules['matplotlib'] = mpl
    sys.modules['matplotlib.pyplot'] = plt


def optimizer_cases(sel):
    """Known answers for the optimiser step and the whole identification loop, from the reference's own classes:
    CustomBFGSREALOptimizer (optimization_lib/optimizers.py:7-111) driven with seeded gradients, and
    PowerSystemOptimization.run() (optimization_lib/ps_optimization.py:150-256) on the IBB model."""
    import torch
    rh._import_reference('torch')
    _stub_matplotlib()
    from src.diffpssi.optimization_lib.optimizers import CustomBFGSREALOptimizer
    name = 'bfgs_reference'
    if not sel or any(s in name for s in sel):
        rng = np.random.RandomState(11)
        for n, tag in ((3, ''), (5, '_n5'), (1, '_n1')):
            P, K = 6, 7
            x0 = rng.uniform(0.5, 3.0, (n, P))
            grads = rng.normal(0, 1, (K, n, P)) * rng.uniform(0.01, 10.0, (1, n, 1))
            params = [torch.tensor(x0[i].reshape(P, 1), dtype=torch.complex128, requires_grad=True) for i in range(n)]
            opt = CustomBFGSREALOptimizer(params, max_step=0.1, decay=0.75, verbose=False)
            traj, steps = [], []
            for k in range(K):
                for i, p_ in enumerate(params):
                    p_.grad = torch.complex(torch.tensor(grads[k, i].reshape(P, 1)), torch.full((P, 1), 0.123 * (k + 1), dtype=torch.float64))
                opt.step()
                if k in (2, 4):
                    opt.decrease_step_size()
                traj.append(np.stack([p_.data.real.numpy().reshape(-1).copy() for p_ in params]))
                steps.append(opt.max_step)
            np.savez_compressed(os.path.join(GOLD, name + tag + '.npz'), x0=x0, grads=grads, traj=np.stack(traj),
                                max_step_after=np.array(steps), decay=0.75, max_step=0.1, decrease_after=np.array([2, 4]),
                                hesse=opt.hesse.numpy())
            print('wrote', name + tag)
    name = 'opt_ibb_reference'
    if not sel or any(s in name for s in sel):
        from src.diffpssi.optimization_lib.ps_optimization import PowerSystemOptimization
        rng = np.random.RandomState(21)
        P, K = 6, 4
        grid = rh.load_example_grid(*GRIDS['ibb'])
        ov = {('gen', 'Gen 1', 'h'): 3.5 * rng.uniform(0.6, 1.6, P),
              ('gen', 'Gen 1', 'x_d_t'): 0.3 * rng.uniform(0.6, 1.6, P),
              ('gen', 'Gen 1', 'd'): rng.uniform(0.3, 2.0, P)}
        sim, leaves = rh.build(grid, P, 'heun', DT, 3.0, [(1, 1.05, 'Bus 1')], ov, requires_grad=True)
        sim.set_record_function(lambda s_: [s_.busses[1].models[0].omega.real, s_.busses[1].models[0].e_q_st.real,
                                            s_.busses[1].models[0].e_d_st.real])
        target = torch.tensor(np.load(os.path.join(rh.REFERENCE_ROOT, 'tests', 'integration', 'data', 'ibb_model.npy'))[:600])
        target = target * torch.ones((P,) + tuple(target.shape))
        keys = list(leaves.keys())
        params = [leaves[k] for k in keys]
        opt = PowerSystemOptimization(sim, target, params_optimizable=params, param_names=['|'.join(k) for k in keys],
                                      decay=0.75, max_step=0.1)
        traj, losses = [], []
        orig_step = opt.optimizer.step

        def spy(**kw):
            orig_step(**kw)
            traj.append(np.stack([p_.data.real.numpy().reshape(-1).copy() for p_ in params]))
        opt.optimizer.step = spy
        orig_loss = opt.loss_function

        def loss_spy(a, b):
            out = orig_loss(a, b)
            losses.append(out.detach().numpy().copy())
            return out
        opt.loss_function = loss_spy
        os.environ.pop('DIFFPSSI_FORCE_OPT_ITERS', None)
        opt.run(max_steps=K)
        extra = {'override/' + '|'.join(k): v for k, v in ov.items()}
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), keys=np.array(['|'.join(k) for k in keys]), traj=np.stack(traj),
                            losses=np.stack(losses), target=target[0].numpy(), decay=0.75, max_step=0.1,
                            final_max_step=opt.optimizer.max_step, **extra)
        print('wrote', name, 'losses', np.stack(losses).min(axis=1))


def wide_grad_case(sel):
    """Gradients on a system with more than 16 generators (the wide stepper / wide adjoint): tiled grid with 20 cells
    (100 buses, 20 machines with SEXS+TGOV1+STAB1), RK4, machine and controller leaves on several machines; reference
    autograd, FD-checked on the first leaf."""
    import torch
    from diffpssi_b200 import grids
    name = 'grad_tiled20_rk4'
    if sel and not any(s in name for s in sel):
        return
    rng = np.random.RandomState(17)
    P, T = 3, 160
    grid = grids.tiled(20)
    ov = {('gen', 'M3', 'h'): 6.175 * rng.uniform(0.8, 1.25, P),
          ('gen', 'M7', 'x_d_t'): 0.3 * rng.uniform(0.8, 1.25, P),
          ('gen', 'M12', 'd'): rng.uniform(0.5, 2.0, P),
          ('gen', 'M18', 'exciter.pt1_lim.gain_pt1'): 100.0 * rng.uniform(0.8, 1.25, P),
          ('gen', 'M0', 'governor.lead_lag.t_1'): 1.0 * rng.uniform(0.8, 1.25, P)}
    events = [(0.1, 0.15, 'A3')]
    chans = ['M%d.omega' % c for c in (0, 3, 7, 12, 18)] + ['M3.e_q_st']

    def run(ovv, grad):
        sim, leaves = rh.build(grid, P, 'rk4', DT, T * DT, events, ovv, requires_grad=grad)
        t, rec, names, snap, rec_t = rh.run_full(sim)
        idx = [names.index(c) for c in chans]
        return leaves, rec_t[:, :, idx], names, snap, t, rec

    leaves, out, names, snap, t, rec = run(ov, True)
    tt = torch.tensor(t)[None, :, None]
    target = 0.002 * torch.sin(2.0 * tt) * torch.ones_like(out.real)
    loss = ((out.real - target) ** 2).mean(dim=(1, 2))
    loss.sum().backward()
    extra = {'override/' + '|'.join(k): v for k, v in ov.items()}
    extra.update({'grad/' + '|'.join(k): v.grad.real.numpy().reshape(-1).copy() for k, v in leaves.items()})
    extra['loss'] = loss.detach().numpy()
    extra['loss_channels'] = np.array(chans)
    extra['target'] = target[0, :, 0].numpy()
    k0 = list(ov.keys())[0]
    h = 1e-6 * np.abs(ov[k0]).max()
    lp = []
    for sgn in (+1, -1):
        ov2 = dict(ov)
        ov2[k0] = ov[k0] + sgn * h
        _, o2, _, _, _, _ = run(ov2, False)
        lp.append(((o2.real - target) ** 2).mean(dim=(1, 2)).numpy())
    fd = (lp[0] - lp[1]) / (2 * h)
    g = extra['grad/' + '|'.join(k0)]
    print('   FD check on %s: max rel diff %.2e' % ('|'.join(k0), np.abs(fd - g).max() / np.abs(g).max()))
    keep = [i for i, n in enumerate(names) if not n.startswith('bus:')]
    save_case(name, t, rec[:, :, keep], [names[i] for i in keep], snap, events, 'rk4', extra, stride=20)


def main():
    sel = sys.argv[1:]
    os.makedirs(os.path.join(GOLD, 'grids'), exist_ok=True)
    for gname, (ex, mod) in GRIDS.items():
        with open(os.path.join(GOLD, 'grids', gname + '.json'), 'w') as f:
            json.dump(rh.load_example_grid(ex, mod), f, indent=1)
    for ref in ('ibb_model', 'ibb_transformer', 'ibb_with_controllers', 'ieee_9bus', 'k2a'):
        shutil.copyfile(os.path.join(rh.REFERENCE_ROOT, 'tests', 'integration', 'data', ref + '.npy'),
                        os.path.join(GOLD, 'ref_' + ref + '.npy'))
        os.chmod(os.path.join(GOLD, 'ref_' + ref + '.npy'), 0o644)
    example_cases(sel)
    batch_cases(sel)
    param_event_cases(sel)
    grad_cases(sel)
    headline_cases(sel)
    ident_range_grad_case(sel)
    optimizer_cases(sel)
    wide_grad_case(sel)


if __name__ == '__main__':
    main()
