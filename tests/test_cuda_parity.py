"""GPU parity tests: the CUDA engine, called through the C ABI (ctypes -> libdps_b200.so), against
  (1) the golden vectors generated from the reference's torch backend (tests/golden/*.npz, tolerance 1e-9 p.u. abs),
  (2) the reference's own golden arrays (tests/golden/ref_*.npy, the reference's tolerance 1e-6),
  (3) the C oracle (oracle/dps_oracle.c) on the same seeded inputs at sizes it finishes in seconds,
  (4) size-independent properties at full batch sizes (batch invariance, tail handling, determinism).
"""
import numpy as np
import pytest
import torch

from diffpssi_b200 import Pss, engine
from tests import helpers
from tests.test_host_pipeline import CASES

pytestmark = pytest.mark.gpu

TOL = 1e-9          # BASELINE.json north_star: trajectories within 1e-9 p.u. absolute over the full simulation


@pytest.mark.parametrize('case,grid', CASES)
def test_golden_trajectories(case, grid):
    sim, gold, _ = helpers.sim_from_gold(case, grid, device='cuda')
    t, res = sim.run()
    assert res.is_cuda and tuple(res.shape) == (gold['traj'].shape[0], len(gold['time']), len(gold['names']))
    mine = res.cpu().numpy()[:, gold['steps'], :]
    err = np.abs(mine - gold['traj']).max()
    assert err < TOL, err
    assert np.array_equal(t, gold['time'])


EXAMPLES = [('ibb', 'ibb_model'), ('ibb_trans', 'ibb_transformer'), ('ibb_wc', 'ibb_with_controllers'),
            ('ieee9', 'ieee_9bus'), ('k2a', 'k2a')]


def example_record_function(grid):
    """The record functions of the reference's example scripts (examples/models/*/*_sim.py)."""
    if grid == 'ibb':
        return lambda s: [s.busses[1].models[0].omega.real, s.busses[1].models[0].e_q_st.real,
                          s.busses[1].models[0].e_d_st.real]
    if grid == 'ibb_trans':
        return lambda s: [s.busses[2].models[0].omega.real, s.busses[2].models[0].e_q_st.real,
                          s.busses[2].models[0].e_d_st.real]
    if grid == 'ibb_wc':
        return lambda s: [s.busses[1].models[0].omega.real]
    if grid == 'ieee9':
        return lambda s: [s.busses[0].models[0].omega.real, s.busses[0].models[0].p_e,
                          s.busses[0].models[0].v_bb.real, s.busses[0].models[0].v_bb.imag,
                          s.busses[1].models[0].omega.real, s.busses[1].models[0].p_e,
                          s.busses[1].voltage.real, s.busses[1].voltage.imag,
                          s.busses[2].models[0].omega.real, s.busses[2].models[0].p_e,
                          s.busses[2].voltage.real, s.busses[2].voltage.imag]
    return lambda s: ([s.busses[i].models[0].delta.real for i in range(4)] +
                      [s.busses[i].models[0].omega.real for i in range(4)] +
                      [s.busses[i].models[0].e_q_st.real for i in range(4)] +
                      [s.busses[i].models[0].e_d_st.real for i in range(4)])


@pytest.mark.parametrize('grid,ref_file', EXAMPLES)
def test_reference_example_scripts(grid, ref_file):
    """The reference's own integration tests (tests/integration/test_example_sims.py), re-stated: same grid dict,
    same event, same Python record function, compared with the reference's golden array at its tolerance (1e-6) and
    with the live torch-backend output at 1e-9."""
    sim, gold, _ = helpers.sim_from_gold(grid + '_heun', grid, device='cuda')
    sim.set_record_function(example_record_function(grid))
    t, res = sim.run()
    ref = np.load(helpers.GOLD + '/ref_%s.npy' % ref_file)
    out = res[0].cpu().numpy()
    assert np.allclose(out, ref, atol=1e-6, rtol=1e-6)
    assert np.abs(out - gold['example_rec']).max() < TOL


def test_parallel_sims_batch_invariance():
    """tests/integration/test_parallel_sims.py: IEEE-9 with parallel_sims=3, every row equals the P=1 golden
    (tolerance 1e-8 there)."""
    sim, gold, _ = helpers.sim_from_gold('ieee9_heun', 'ieee9', device='cuda', P=3)
    sim.set_record_function(example_record_function('ieee9'))
    _, res = sim.run()
    ref = np.load(helpers.GOLD + '/ref_ieee_9bus.npy')
    for row in range(3):
        assert np.allclose(res[row].cpu().numpy(), ref, atol=1e-8, rtol=1e-8)
    assert torch.equal(res[0], res[1]) and torch.equal(res[0], res[2])


def _random_overrides(grid, P, rng, lo, hi):
    ov = {}
    nominal = {'k2a': {'G1': (6.5, 0.3), 'G2': (6.5, 0.3), 'G3': (6.175, 0.3), 'G4': (6.175, 0.3)},
               'ieee9': {'G1': (9.55, 0.15048), 'G2': (3.92, 0.230016), 'G3': (2.766544, 0.232064)},
               'ibb': {'Gen 1': (3.5, 0.3)}, 'ibb_wc': {'G1': (3.5, 0.3)}}[grid]
    for g, (h, xdt) in nominal.items():
        ov[(g, 'h')] = h * rng.uniform(lo, hi, P)
        ov[(g, 'd')] = rng.uniform(0.0, 2.0, P)
        ov[(g, 'x_d_t')] = xdt * rng.uniform(lo, hi, P)
    return ov


@pytest.mark.parametrize('grid,solver,P,events,T', [
    ('k2a', 'heun', 200, [(0.1, 0.2, 'B1')], 300),
    ('k2a', 'rk4', 70, [(0.1, 0.15, 'B8')], 200),
    ('ieee9', 'euler', 97, [(0.1, 0.15, 'B8')], 300),
    ('ibb_wc', 'heun', 33, [(0.1, 0.15, 'B2')], 400),
    ('ibb', 'heun', 1, [(0.1, 0.15, 'Bus 1')], 400),
])
def test_against_oracle_on_seeded_batches(grid, solver, P, events, T):
    """CUDA path vs the C oracle on identical inputs (the package's own prologue output): every state and algebraic of
    every generator at every step, ragged batch sizes (not multiples of 32), faults at generator and load buses."""
    rng = np.random.RandomState(1234)
    g = helpers.load_grid(grid)
    sim = Pss(parallel_sims=P, sim_time=T * 0.005, time_step=0.005, solver=solver, grid_data=g, verbose=False)
    for ev in events:
        sim.add_sc_event(*ev)
    if P > 1:
        helpers.apply_overrides(sim, _random_overrides(grid, P, rng, 0.9, 1.1))
    init = sim.initialize()
    names = []
    for gen in init.gens:
        names += ['%s.%s' % (gen.name, f) for f in helpers.oracle.FIELD_NAMES]
    sim.set_record_channels(names)
    _, res = sim.run()
    ref = helpers.run_oracle_for(sim, sim.last_init)
    want = ref['full'].transpose(3, 0, 1, 2).reshape(P, len(sim.time), -1)      # [P, T, G*20]
    err = np.abs(res.cpu().numpy() - want).max()
    assert err < TOL, err


def test_recorded_bus_voltages_against_oracle():
    """Voltages of non-generator buses (extra observed rows of the reduced impedance matrix), incl. the faulted bus."""
    g = helpers.load_grid('k2a')
    P = 40
    sim = Pss(parallel_sims=P, sim_time=1.0, time_step=0.005, solver='heun', grid_data=g, verbose=False)
    sim.add_sc_event(0.2, 0.3, 'B8')
    names = ['bus:%s.%s' % (b.name, p) for b in sim.busses for p in ('v_re', 'v_im')]
    sim.set_record_channels(names)
    _, res = sim.run()
    ref = helpers.run_oracle_for(sim, sim.last_init)
    want = ref['busv'].transpose(3, 0, 1, 2).reshape(P, len(sim.time), -1)
    assert np.abs(res.cpu().numpy() - want).max() < TOL


def test_determinism_and_rerun_after_reset():
    sim, gold, _ = helpers.sim_from_gold('k2a_heun', 'k2a', device='cuda', P=64, sim_time=1.5)
    _, a = sim.run()
    sim.reset()
    _, b = sim.run()
    assert torch.equal(a, b)


def test_fp64_probe_runs():
    tf, ms = engine.fp64_peak_probe(2048)
    assert 5.0 < tf < 80.0, tf


def test_streamed_recorder_equals_plain_run():
    """run(host_out=...) cuts the loop into step ranges (dps_forward_range) and streams the recorder to pinned host
    memory; the trajectories must be bit-identical to the single-launch run."""
    sim, gold, _ = helpers.sim_from_gold('k2a_heun', 'k2a', device='cuda', P=70, sim_time=2.0)
    _, a = sim.run()
    sim.reset()
    host = torch.empty((len(gold['names']), len(sim.time), 70), dtype=torch.float64).pin_memory()
    _, b = sim.run(host_out=host, stream_chunks=7)
    assert not b.is_cuda and tuple(b.shape) == tuple(a.shape)
    assert torch.equal(a.cpu(), b)


def test_device_math():
    """The branch-free sincos / division / sqrt the kernels use, against numpy (libm): <= 2 ulp."""
    rng = np.random.RandomState(3)
    a = np.concatenate([rng.uniform(-200.0, 200.0, 200000), rng.uniform(-1e-3, 1e-3, 1000), rng.uniform(-3e4, 3e4, 20000),
                        np.array([0.0, 1e-300, np.pi / 2, -np.pi / 4, 1.0])])
    b = np.concatenate([rng.uniform(0.2, 3.0, a.size - 5) * rng.choice([-1.0, 1.0], a.size - 5),
                        np.array([1.0, 3.0, 1e-3, 7.0, 1e6])])
    out = engine.math_probe(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    eps = np.finfo(np.float64).eps

    def ulps(x, ref, floor=0.0):
        return np.abs(x - ref) / (eps * np.maximum(np.abs(ref), floor))

    assert ulps(out[0], np.sin(a), 1e-3).max() < 4 and ulps(out[1], np.cos(a), 1e-3).max() < 4
    assert np.abs(out[0] - np.sin(a)).max() < 1e-15 and np.abs(out[1] - np.cos(a)).max() < 1e-15
    assert ulps(out[2], a / b, 1e-300).max() <= 1.0
    assert ulps(out[3], np.sqrt(np.abs(a)), 1e-300).max() <= 1.0
    assert ulps(out[4], 1.0 / b).max() <= 1.0


def _ieee39_sim(P, rows_params, device='cuda', T=2000):
    from diffpssi_b200 import grids
    sim = Pss(parallel_sims=P, sim_time=T * 0.005, time_step=0.005, solver='heun', grid_data=grids.ieee39(), verbose=False,
              device=device)
    sim.add_sc_event(1.0, 1.1, 'B16')
    helpers.apply_overrides(sim, rows_params)
    sim.set_record_channels(['G%d.omega' % i for i in range(1, 11)])
    return sim


def test_full_size_ieee39_65536():
    """BASELINE config 3 at full size (65 536 trajectories, 2000 Heun steps): 32 distinct parameter rows tiled 2048 times.
    Size-independent properties: every tile is bit-identical to the first (rows are independent, no cross-talk, no
    tail/indexing error anywhere in the batch) and the first tile equals a 32-trajectory run, which in turn matches the
    C oracle on the same inputs to 1e-9."""
    from diffpssi_b200 import grids
    rng = np.random.RandomState(5)
    n, tile = 32, 2048
    small = {}
    for g, (h, xdt) in grids.GENERATOR_NOMINALS['ieee39'].items():
        small[(g, 'h')] = h * rng.uniform(0.8, 1.25, n)
        small[(g, 'd')] = rng.uniform(0.0, 2.0, n)
        small[(g, 'x_d_t')] = xdt * rng.uniform(0.9, 1.1, n)
    big = {k: np.tile(v, tile) for k, v in small.items()}
    sim_big = _ieee39_sim(n * tile, big)
    _, res_big = sim_big.run()
    assert tuple(res_big.shape) == (n * tile, 2000, 10) and bool(torch.isfinite(res_big).all())
    first = res_big[:n]
    for t in (1, 777, tile - 1):
        assert torch.equal(res_big[t * n:(t + 1) * n], first)
    chk = res_big.reshape(tile, n, 2000, 10)
    assert torch.equal(chk.amax(dim=0), chk.amin(dim=0))                 # all 2048 tiles identical, everywhere
    sim_small = _ieee39_sim(n, small)
    _, res_small = sim_small.run()
    assert torch.equal(res_small, first)
    ref = helpers.run_oracle_for(sim_small, sim_small.last_init, rec_channels=[(0, g, 0) for g in range(10)], full=False)
    want = ref['rec'].transpose(2, 1, 0)
    assert np.abs(res_small.cpu().numpy() - want).max() < TOL


def test_recorder_specialisations_agree():
    """The same IEEE-39 batch recorded three ways -- omega of every generator (compile-time recorder mask), omega + delta of
    every generator (regular layout, run-time mask) and an irregular channel list (per-thread channel table) -- must give
    bit-identical omega trajectories: the recorder path must not influence the integration."""
    from diffpssi_b200 import grids
    rng = np.random.RandomState(11)
    P = 77                                                   # ragged: 25 full warps of 3 trajectories + 2
    ov = {}
    for g, (h, xdt) in grids.GENERATOR_NOMINALS['ieee39'].items():
        ov[(g, 'h')] = h * rng.uniform(0.8, 1.25, P)
        ov[(g, 'd')] = rng.uniform(0.0, 2.0, P)
    gens = ['G%d' % i for i in range(1, 11)]
    layouts = {
        'omega': ['%s.omega' % g for g in gens],
        'regular': ['%s.omega' % g for g in gens] + ['%s.delta' % g for g in gens],
        'irregular': ['G3.p_e'] + ['%s.omega' % g for g in gens] + ['G7.delta', 'bus:B16.v_re'],
    }
    out = {}
    for name, chans in layouts.items():
        sim = _ieee39_sim(P, ov)
        sim.set_record_channels(chans)
        _, res = sim.run()
        out[name] = res[:, :, [chans.index('%s.omega' % g) for g in gens]].clone()
        assert bool(torch.isfinite(res).all())
    assert torch.equal(out['omega'], out['regular'])
    assert torch.equal(out['omega'], out['irregular'])


def test_pair_kernel_matches_single_trajectory_kernel(monkeypatch):
    """dps_step_pair_kernel (two trajectories per thread, used for large batches of the all-controllers / shared-network /
    omega-recorder configuration) against dps_step_kernel on the same ragged batch, and against the C oracle."""
    from diffpssi_b200 import grids
    rng = np.random.RandomState(21)
    P = 77                                                   # ragged: 12 full warps of 6 trajectories + 5
    ov = {}
    for g, (h, xdt) in grids.GENERATOR_NOMINALS['ieee39'].items():
        ov[(g, 'h')] = h * rng.uniform(0.8, 1.25, P)
        ov[(g, 'd')] = rng.uniform(0.0, 2.0, P)
        ov[(g, 'x_d_t')] = xdt * rng.uniform(0.9, 1.1, P)
    out, regs = {}, {}
    for name, min_sims in (('pair', '0'), ('single', '999999999')):
        monkeypatch.setenv('DPS_PAIR_MIN_SIMS', min_sims)
        sim = _ieee39_sim(P, ov, T=600)
        _, res = sim.run()
        out[name] = res.clone()
        regs[name] = sim.last_plan.info()['regs_per_thread']
    assert regs['pair'] != regs['single']                    # the two launches really used different kernels
    assert float((out['pair'] - out['single']).abs().max()) < 1e-12
    monkeypatch.setenv('DPS_PAIR_MIN_SIMS', '0')
    sim = _ieee39_sim(P, ov, T=600)
    _, res = sim.run()
    # step ranges (dps_forward_range: state carried through x_out, recorder pointers offset by step_begin)
    sim.reset()
    host = torch.empty((10, 600, P), dtype=torch.float64).pin_memory()
    _, streamed = sim.run(host_out=host, stream_chunks=5)
    assert torch.equal(streamed, res.cpu())
    ref = helpers.run_oracle_for(sim, sim.last_init, rec_channels=[(0, g, 0) for g in range(10)], full=False)
    assert np.abs(res.cpu().numpy() - ref['rec'].transpose(2, 1, 0)).max() < TOL


def test_full_size_k2a_4096():
    """BASELINE config 2 (K2A, 4096 trajectories, 2000 Heun steps, per-sim H/D/X_d' in the stable range): 24 sampled
    rows re-run as a small batch are bit-identical, and match the C oracle to 1e-9."""
    rng = np.random.RandomState(6)
    P = 4096
    g = helpers.load_grid('k2a')
    ov = _random_overrides('k2a', P, rng, 0.9, 1.1)

    def build(Pn, params):
        s = Pss(parallel_sims=Pn, sim_time=10, time_step=0.005, solver='heun', grid_data=g, verbose=False)
        s.add_sc_event(1, 1.1, 'B1')
        helpers.apply_overrides(s, params)
        s.set_record_channels(['G%d.%s' % (i, f) for f in ('delta', 'omega', 'e_q_st', 'e_d_st') for i in (1, 2, 3, 4)])
        return s

    big = build(P, ov)
    _, res = big.run()
    assert bool(torch.isfinite(res).all())
    rows = np.sort(rng.choice(P, 24, replace=False))
    small = build(24, {k: v[rows] for k, v in ov.items()})
    _, res_s = small.run()
    assert torch.equal(res[torch.from_numpy(rows).cuda()], res_s)
    names = [('G%d' % i, f) for f in ('delta', 'omega', 'e_q_st', 'e_d_st') for i in (1, 2, 3, 4)]
    rec = [(0, int(n_[0][1]) - 1, helpers.oracle.FIELD_NAMES.index(n_[1])) for n_ in names]
    ref = helpers.run_oracle_for(small, small.last_init, rec_channels=rec, full=False)
    assert np.abs(res_s.cpu().numpy() - ref['rec'].transpose(2, 1, 0)).max() < TOL


@pytest.mark.parametrize('n_cells,P,T,solver', [(20, 37, 60, 'rk4'), (40, 9, 40, 'heun'), (100, 6, 30, 'rk4')])
def test_wide_stepper_tiled_grid(n_cells, P, T, solver):
    """BASELINE config 5 family: tiled 5-bus cells with all controllers, 17..128 generators (dps_step_wide_kernel), RK4 /
    Heun, per-sim H and X_d', a fault at a load bus, recorded generator and load-bus quantities -- vs the C oracle
    (full n x n complex network, two dense inversions per active fault step) on the same inputs."""
    from tests.test_host_pipeline import _tiled_case
    sim, rec = _tiled_case(n_cells, P, T, solver, 'cuda')
    _, res = sim.run()
    info = sim.last_plan.info()
    assert info['g_template'] == n_cells and info['block_threads'] == 512
    ref = helpers.run_oracle_for(sim, sim.last_init, rec_channels=rec, full=False)
    assert np.abs(res.cpu().numpy() - ref['rec'].transpose(2, 1, 0)).max() < TOL


# ----------------------------------------------------------------------------------------------------------------------
# parity ON the benchmarked configurations themselves (task 1)
# ----------------------------------------------------------------------------------------------------------------------
def _per_row_parity(mine, want, delta_rows):
    """mine/want [P,T,C].  Rows that keep synchronism (|delta| < 10 rad) must agree to 1e-9 absolute; pole-slipping rows
    amplify round-off beyond that even between the reference's own two backends, so every row
    must agree to 1e-6 relative.  Returns the number of stable rows."""
    err = np.abs(mine - want).max(axis=(1, 2))
    stable = np.abs(delta_rows).max(axis=(1, 2)) < 10.0
    assert (err[stable] < TOL).all(), (err[stable].max(), int(np.argmax(err)))
    scale = np.abs(want).max(axis=(1, 2))
    assert (err / scale < 1e-6).all(), (err / scale).max()
    return int(stable.sum())


def test_bench_parameter_draw_against_oracle():
    """IEEE-39 with the BENCH's own parameter draw (bench.parameter_batch: H, X_d' ~ U(0.5,2) x nominal, D ~ U(0,2), seed
    0), rows 0..511, full 2000 Heun steps, the bench's fault and recorder (+ delta): CUDA vs the C oracle, per row."""
    import bench
    P = 512
    batch = bench.parameter_batch(bench.P_GLOBAL)
    sim = bench.build_sim(P, 'cuda', (0, P), batch)
    sim.set_record_channels(['G%d.omega' % i for i in range(1, 11)] + ['G%d.delta' % i for i in range(1, 11)])
    _, res = sim.run()
    ref = helpers.run_oracle_for(sim, sim.last_init, rec_channels=[(0, g, 0) for g in range(10)] + [(0, g, 1) for g in range(10)],
                                 full=False)
    want = ref['rec'].transpose(2, 1, 0)
    n_stable = _per_row_parity(res.cpu().numpy(), want, want[:, :, 10:])
    assert n_stable >= P // 2


def test_bench_workload_full_size_sampled_rows():
    """The bench workload exactly as bench.py builds it (65 536 trajectories of the seed-0 draw, omega recorder -> the
    two-trajectories-per-thread kernel): 256 sampled rows (+ the four rows the reference itself ran for
    tests/golden/ieee39_batch_heun.npz) against the C oracle, and against that golden."""
    import bench
    rng = np.random.RandomState(12)
    P = bench.P_GLOBAL
    batch = bench.parameter_batch(P)
    sim = bench.build_sim(P, 'cuda', (0, P), batch)
    _, res = sim.run()
    assert sim.last_plan.info()['sims_per_block'] == 24                 # dps_step_pair_kernel took the launch
    assert tuple(res.shape) == (P, bench.T_STEPS, 10)
    rows = np.concatenate([np.arange(4), np.sort(rng.choice(np.arange(4, P), 256, replace=False))])
    mine = res[torch.from_numpy(rows).cuda()].cpu().numpy()
    del res
    small = bench.build_sim(len(rows), 'cuda', (0, len(rows)), {k: v[rows] for k, v in batch.items()})
    small.set_record_channels(['G%d.omega' % i for i in range(1, 11)] + ['G%d.delta' % i for i in range(1, 11)])
    _, res_s = small.run()
    ref = helpers.run_oracle_for(small, small.last_init,
                                 rec_channels=[(0, g, 0) for g in range(10)] + [(0, g, 1) for g in range(10)], full=False)
    want = ref['rec'].transpose(2, 1, 0)
    _per_row_parity(mine, want[:, :, :10], want[:, :, 10:])
    gold = helpers.load_gold('ieee39_batch_heun')
    names = [str(s) for s in gold['names']]
    cols = [names.index('G%d.omega' % i) for i in range(1, 11)]
    assert np.abs(mine[:4][:, gold['steps'], :] - gold['traj'][:, :, cols]).max() < TOL


def test_config5_full_length_sampled_rows():
    """BASELINE config 5 as tools/bench_configs.py times it (tiled 500-bus / 100-generator grid, RK4, 8192 trajectories,
    the full T = 2000 steps, per-sim H / D / X_d' from the 0.5x..2x range): sampled rows re-run as a small batch are
    bit-identical, and match the C oracle (full 500 x 500 complex network) per row."""
    from diffpssi_b200 import grids
    P, T = 8192, 2000
    rng = np.random.RandomState(0)
    ov = {}
    for c in range(100):
        ov[('M%d' % c, 'h')] = (6.5 if c % 2 == 0 else 6.175) * rng.uniform(0.5, 2, P)
        ov[('M%d' % c, 'd')] = rng.uniform(0, 2, P)
        ov[('M%d' % c, 'x_d_t')] = 0.3 * rng.uniform(0.5, 2, P)

    def build(Pn, params, chans):
        s = Pss(parallel_sims=Pn, sim_time=T * 0.005, time_step=0.005, solver='rk4', grid_data=grids.tiled(100), verbose=False)
        s.add_sc_event(1.0, 1.1, 'A3')
        helpers.apply_overrides(s, params)
        s.set_record_channels(chans)
        return s

    omega = ['M%d.omega' % c for c in range(100)]
    big = build(P, ov, omega)
    _, res = big.run()
    assert big.last_plan.info()['block_threads'] == 512                  # dps_step_wide_kernel
    assert bool(torch.isfinite(res).all())
    rows = np.array([0, 4097, 8191])
    small = build(len(rows), {k: v[rows] for k, v in ov.items()}, omega + ['M%d.delta' % c for c in range(100)])
    _, res_s = small.run()
    assert torch.equal(res[torch.from_numpy(rows).cuda()], res_s[:, :, :100])
    ref = helpers.run_oracle_for(small, small.last_init,
                                 rec_channels=[(0, c, 0) for c in range(100)] + [(0, c, 1) for c in range(100)], full=False)
    want = ref['rec'].transpose(2, 1, 0)
    _per_row_parity(res_s.cpu().numpy(), want, want[:, :, 100:])


def test_every_bus_voltage_of_ieee39_can_be_recorded():
    """static_models.py:153-163 keeps the voltage of every bus; recording all 39 of IEEE-39 needs 29 extra observed rows
    of the reduced impedance matrix (the round-1 cap was 24)."""
    sim = _ieee39_sim(9, {}, T=300)
    names = ['bus:%s.%s' % (b.name, p) for b in sim.busses for p in ('v_re', 'v_im')]
    sim.set_record_channels(names)
    _, res = sim.run()
    ref = helpers.run_oracle_for(sim, sim.last_init)
    want = ref['busv'].transpose(3, 0, 1, 2).reshape(9, len(sim.time), -1)
    assert np.abs(res.cpu().numpy() - want).max() < TOL


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_simulation_on_a_non_current_device():
    """Pss(device='cuda:1') while device 0 is current: every C-ABI call must run on the device that holds the buffers."""
    torch.cuda.set_device(0)
    sim, gold, _ = helpers.sim_from_gold('k2a_heun', 'k2a', device='cuda:1', P=33, sim_time=1.5)
    _, a = sim.run()
    assert a.device == torch.device('cuda:1') and torch.cuda.current_device() == 0
    sim0, _, _ = helpers.sim_from_gold('k2a_heun', 'k2a', device='cuda:0', P=33, sim_time=1.5)
    _, b = sim0.run()
    assert torch.equal(a.cpu(), b.cpu())


def test_host_prologue_equals_device_prologue():
    """Pss(prologue_device='cpu'): load flow / initialisation on the host, packed inputs uploaded, same kernels."""
    out = {}
    for pdev in ('cpu', None):
        g = helpers.load_grid('k2a')
        sim = Pss(parallel_sims=40, sim_time=1.5, time_step=0.005, solver='heun', grid_data=g, verbose=False, device='cuda',
                  prologue_device=pdev)
        sim.add_sc_event(0.2, 0.3, 'B8')
        rng = np.random.RandomState(2)
        helpers.apply_overrides(sim, _random_overrides('k2a', 40, rng, 0.9, 1.1))
        sim.set_record_channels(['G%d.%s' % (i, f) for f in ('delta', 'omega', 'p_e', 'e_fd') for i in (1, 2, 3, 4)] + ['bus:B8.v_re'])
        _, res = sim.run()
        assert res.is_cuda
        out[pdev] = res.cpu().numpy()
    assert np.abs(out['cpu'] - out[None]).max() < 1e-12


@pytest.mark.parametrize('pair_min', ['0', '999999999'])
def test_work_queue_matches_static_launch(monkeypatch, pair_min):
    """The (block group, step chunk) work queue (csrc/dps_internal.h: q_*) only reschedules the same arithmetic: a batch that
    makes ~2.4 waves of blocks must come out bit-identical with the queue on (default) and off, for both stepping kernels,
    with a fault inside one chunk and a chunk boundary inside the fault."""
    from diffpssi_b200 import grids
    rng = np.random.RandomState(31)
    P = 148 * 3 * 12 * 2 + 517                    # > 2 waves of dps_step_kernel blocks, ragged
    ov = {}
    for g, (h, xdt) in grids.GENERATOR_NOMINALS['ieee39'].items():
        ov[(g, 'h')] = h * rng.uniform(0.6, 1.8, P)
        ov[(g, 'd')] = rng.uniform(0.0, 2.0, P)
        ov[(g, 'x_d_t')] = xdt * rng.uniform(0.6, 1.8, P)
    monkeypatch.setenv('DPS_PAIR_MIN_SIMS', pair_min)
    out = {}
    for rounds in ('0', '12', '40'):
        monkeypatch.setenv('DPS_QUEUE_ROUNDS', rounds)
        sim = _ieee39_sim(P, ov, T=500)
        sim.sc_events = []
        sim.add_sc_event(0.9, 1.3, 'B16')
        _, res = sim.run()
        out[rounds] = res.clone()
        assert bool(torch.isfinite(res).all())
    assert torch.equal(out['0'], out['12']) and torch.equal(out['0'], out['40'])


def test_generic_pair_kernel_matches_single_trajectory_kernel(monkeypatch):
    """The general build of dps_step_pair_kernel (per-generator controller constants from shared memory, run-time recorder
    mask) on bench.py's ``--variant generic`` grid (distinct controller settings per machine, omega + delta recorder)
    against dps_step_kernel on the same ragged batch, and against the C oracle."""
    import bench
    rng = np.random.RandomState(23)
    P = 83
    batch = {k: v[:P] for k, v in bench.parameter_batch(4096).items()}
    out = {}
    for name, min_sims in (('pair', '0'), ('single', '999999999')):
        monkeypatch.setenv('DPS_PAIR_MIN_SIMS', min_sims)
        sim = bench.build_sim(P, 'cuda', (0, P), batch, variant='generic')
        _, res = sim.run()
        out[name] = res.clone()
        info = sim.last_plan.info()
        assert (info['regs_per_thread'] > 200) == (name == 'pair'), info
    assert tuple(out['pair'].shape) == (P, bench.T_STEPS, 20)
    assert float((out['pair'] - out['single']).abs().max()) < 1e-12
    ref = helpers.run_oracle_for(sim, sim.last_init, rec_channels=[(0, g, 0) for g in range(10)] + [(0, g, 1) for g in range(10)],
                                 full=False)
    want = ref['rec'].transpose(2, 1, 0)
    _per_row_parity(out['pair'].cpu().numpy(), want, want[:, :, 10:])


@pytest.mark.parametrize('pair_min', ['0', '999999999'])
def test_work_queue_at_small_batch(monkeypatch, pair_min):
    """The queue forced at a CI batch size (DPS_QUEUE_SLOTS pretends the GPU holds 2 blocks): many chunks per group, ragged
    last group, faults crossing chunk boundaries, and dps_forward_range launches on top (streamed recorder)."""
    from diffpssi_b200 import grids
    rng = np.random.RandomState(41)
    P = 131
    ov = {(g, 'h'): h * rng.uniform(0.7, 1.5, P) for g, (h, _x) in grids.GENERATOR_NOMINALS['ieee39'].items()}
    monkeypatch.setenv('DPS_PAIR_MIN_SIMS', pair_min)
    out = {}
    for slots in (None, '2'):
        if slots:
            monkeypatch.setenv('DPS_QUEUE_SLOTS', slots)
        else:
            monkeypatch.delenv('DPS_QUEUE_SLOTS', raising=False)
        sim = _ieee39_sim(P, ov, T=700)
        sim.sc_events = []
        sim.add_sc_event(0.3, 0.7, 'B16')
        sim.add_sc_event(1.5, 1.6, 'B4')
        _, res = sim.run()
        out[slots] = res.clone()
        if slots:
            sim.reset()
            host = torch.empty((10, 700, P), dtype=torch.float64).pin_memory()
            _, streamed = sim.run(host_out=host, stream_chunks=3)
            assert torch.equal(streamed, res.cpu())
    assert torch.equal(out[None], out['2'])


@pytest.mark.parametrize('n_faults', [12, 16])
def test_many_fault_configurations_against_oracle(n_faults):
    """More distinct fault configurations (17 > the former cap of 16) and more simultaneously faulted buses (16 > the former 8)
    than rounds 1-2a accepted.  With the reference's multi-event semantics (simulator.py:524-531: a later inactive event
    wipes the faults of earlier ones) a growing set needs nested events listed latest-start first; the fault re-reduction
    applies one rank-1 update per faulted bus, the oracle inverts twice per active step."""
    from diffpssi_b200 import grids
    P = 5
    sim = Pss(parallel_sims=P, sim_time=0.4, time_step=0.005, solver='heun', grid_data=grids.ieee39(), verbose=False)
    buses = ['B%d' % b for b in range(3, 3 + n_faults)]
    for k in reversed(range(n_faults)):                      # event k starts at 0.102 + 0.01 k, all end at 0.3
        sim.add_sc_event(0.102 + 0.01 * k, 0.3, buses[k])
    rng = np.random.RandomState(8)
    ov = {}
    for g, (h, x) in grids.GENERATOR_NOMINALS['ieee39'].items():
        ov[(g, 'h')] = h * rng.uniform(0.9, 1.1, P)
    helpers.apply_overrides(sim, ov)
    sim.set_record_channels(['G%d.%s' % (i, f) for i in range(1, 11) for f in ('omega', 'delta', 'e_q_st', 'p_e')])
    _, res = sim.run()
    assert sim.last_plan.topo['n_cfg'] == n_faults + 1
    ref = helpers.run_oracle_for(sim, sim.last_init)
    idx = [helpers.oracle.FIELD_NAMES.index(f) for f in ('omega', 'delta', 'e_q_st', 'p_e')]
    want = ref['full'][:, :, idx, :].transpose(3, 0, 1, 2).reshape(P, len(sim.time), -1)
    err = np.abs(res.cpu().numpy() - want).max()
    assert err < TOL, err
