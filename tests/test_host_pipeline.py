"""CPU checks of the host side of the package (grid building, torch prologue, recorder tracing, planning/packing) and
of the device model code compiled for the host (tests/host_emul/emul.cpp), against the golden vectors produced by the
reference's torch backend.  No GPU, no oracle import in the product: see tests/helpers.py."""
import json
import os

import numpy as np
import pytest
import torch

from diffpssi_b200 import Pss, planning, prologue, recording
from tests import helpers

CASES = [(g + '_' + s, g) for g in ('ibb', 'ibb_trans', 'ibb_wc', 'ieee9', 'k2a') for s in ('heun', 'euler', 'rk4')]
CASES += [('k2a_batch_heun', 'k2a'), ('ieee9_persimz_heun', 'ieee9'), ('ieee9_two_events_heun', 'ieee9'),
          ('k2a_two_events_rk4', 'k2a'), ('ibb_param_event_heun', 'ibb'), ('k2a_param_event_rk4', 'k2a')]
# reference outputs on the grids bench.py / BASELINE.json name (IEEE 39-bus incl. rows of the bench's own parameter draw,
# the tiled grid of config 5 at 20 and at the full 100 generators)
CASES += [('ieee39_heun', 'ieee39'), ('ieee39_batch_heun', 'ieee39'), ('tiled20_rk4', 'tiled20'), ('tiled100_rk4', 'tiled100')]


@pytest.mark.parametrize('case,grid', CASES)
def test_prologue_matches_reference_initialize(case, grid):
    """x0, held constants and Z0 after initialize() equal the reference's (simulator.py:438-450) to round-off."""
    sim, gold, _ = helpers.sim_from_gold(case, grid, device='cpu')
    init = sim.initialize()
    z0 = init.Z0.numpy()
    assert np.abs(z0 - gold['init/Z0'][:z0.shape[0]]).max() < 1e-13
    for g, x0, sl in zip(init.gens, init.x0, init.slots):
        ref = gold['init/gen:%s.x0' % g.name]                                    # [P, s]
        mine = np.stack([np.broadcast_to(v.detach().numpy().reshape(-1), (ref.shape[0],)) for v in x0], axis=1)
        assert np.abs(mine - ref).max() < 1e-13
        assert np.abs(sl['pm0'].detach().numpy().reshape(-1) - gold['init/gen:%s.p_m' % g.name]).max() < 1e-13
        assert np.abs(sl['efd0'].detach().numpy().reshape(-1) - gold['init/gen:%s.e_fd' % g.name]).max() < 1e-13


@pytest.mark.parametrize('case,grid', CASES)
def test_emulated_pipeline_matches_reference(case, grid):
    """prologue -> planning -> (host build of the device code) reproduces every state, algebraic and bus voltage of the
    reference torch path to 1e-9."""
    sim, gold, _ = helpers.sim_from_gold(case, grid, device='cpu')
    result, _ = helpers.run_emulated(sim)
    mine = result[:, gold['steps'], :]
    err = np.abs(mine - gold['traj']).max()
    assert err < 1e-9, err


def test_event_schedule_quirks():
    """Fault timing on the np.arange grid: (1, 1.05) -> steps 201..210, (1, 1.1) -> 201..220, and the
    multi-event `else` cancellation of simulator.py:524-531."""
    from diffpssi_b200.components import ScEvent
    t = np.arange(0, 10, 0.005)
    cfg, sets = planning.event_schedule(t, [ScEvent(1, 1.05, 3)])
    assert list(np.nonzero(cfg)[0]) == list(range(201, 211)) and sets == [(), (3,)]
    cfg, sets = planning.event_schedule(t, [ScEvent(1, 1.1, 0)])
    assert list(np.nonzero(cfg)[0]) == list(range(201, 221))
    # later inactive event cancels an earlier active one; later active one survives
    cfg, sets = planning.event_schedule(t, [ScEvent(0.5, 0.6, 7), ScEvent(0.55, 0.7, 4)])
    active = {k: sets[c] for k, c in enumerate(cfg) if c}
    assert all(active[k] == (4, 7) for k in range(111, 121))       # both active: faults at 7 then 4
    # first inactive (wipes), second active; t[140] = 0.7000000000000001 > 0.7 on the float grid, so 139 is the last
    assert all(active[k] == (4,) for k in range(121, 140))
    assert all(k not in active for k in range(101, 111))           # first active, second inactive -> wiped
    assert min(active) == 111 and max(active) == 139


def test_record_function_tracing():
    grid = helpers.load_grid('ieee9')
    sim = Pss(parallel_sims=2, sim_time=0.1, time_step=0.005, solver='heun', grid_data=grid, verbose=False, device='cpu')

    def rec(s):          # reference examples/models/ieee_9bus/ieee_9bus_sim.py:20-36
        return [s.busses[0].models[0].omega.real, s.busses[0].models[0].p_e, s.busses[0].models[0].v_bb.real,
                s.busses[1].voltage.imag, s.busses[7].voltage.real,
                torch.abs(s.busses[7].voltage) * 2.0 + s.busses[2].models[0].h.real]

    init = sim.initialize()
    rp = recording.trace(sim, init.gens, rec)
    assert rp.channel_keys[:5] == [('gen', 0, 0), ('gen', 0, 13), ('gen', 0, 16), ('bus', 1, 1), ('bus', 7, 0)]
    assert ('bus', 7, 1) in rp.channel_keys and not rp.direct
    obs, chans, rows = planning.observed_rows(init, rp.channel_keys)
    assert obs == [0, 1, 2, 7]                       # three generator buses + the recorded load bus B8
    assert chans[3] == (0, 1, 17) and chans[4] == (1, 3, 0)
    fake = torch.arange(len(rp.channel_keys) * 3 * 2, dtype=torch.float64).reshape(len(rp.channel_keys), 3, 2)
    out = rp.assemble(fake)
    assert out.shape == (2, 3, 6)
    k_re, k_im = rp.channel_keys.index(('bus', 7, 0)), rp.channel_keys.index(('bus', 7, 1))
    expect = torch.sqrt(fake[k_re] ** 2 + fake[k_im] ** 2) * 2.0 + 2.766544
    assert torch.allclose(out[:, :, 5], expect.T)


def test_uniform_collapse_and_vary_classification():
    sim, gold, _ = helpers.sim_from_gold('k2a_batch_heun', 'k2a', device='cpu')
    init = sim.initialize()
    assert init.Z0.shape[0] == 1                     # H, D, X_d', PSS/AVR gains do not enter Y: network shared
    upar, vpar, slot_vary = planning.pack_slots(init, sim.parallel_sims, 'cpu')
    names = prologue.SLOTS
    vary = {(i // len(names), names[i % len(names)]) for i, v in enumerate(slot_vary) if v >= 0}
    assert (0, 'inv_2h') in vary and (0, 'd') in vary and (0, 'xd_xdt') in vary and (0, 'xdt_xdst') in vary
    assert (0, 's_kw') in vary and (1, 'a_k') in vary and (0, 'efd0') in vary
    assert (0, 'inv_tq0t') not in vary and (3, 's_kw') not in vary
    sim2, _, _ = helpers.sim_from_gold('ieee9_persimz_heun', 'ieee9', device='cpu')
    assert sim2.initialize().Z0.shape[0] == sim2.parallel_sims        # x_d'' enters Y: per-sim network


def test_backend_argument():
    with pytest.raises(ValueError):
        Pss(parallel_sims=1, sim_time=1, time_step=0.005, solver='heun', backend='numpy')
    with pytest.raises(KeyError):
        Pss(parallel_sims=1, sim_time=1, time_step=0.005, solver='dopri5')


def _tiled_case(n_cells, P, T, solver, device):
    from diffpssi_b200 import grids
    sim = Pss(parallel_sims=P, sim_time=T * 0.005, time_step=0.005, solver=solver, grid_data=grids.tiled(n_cells),
              verbose=False, device=device)
    sim.add_sc_event(0.05, 0.06, 'A3')                     # fault at a load bus, two active steps
    rng = np.random.RandomState(0)
    ov = {}
    for c in range(n_cells):
        ov[('M%d' % c, 'h')] = (6.5 if c % 2 == 0 else 6.175) * rng.uniform(0.9, 1.1, P)
        ov[('M%d' % c, 'x_d_t')] = 0.3 * rng.uniform(0.9, 1.1, P)
    helpers.apply_overrides(sim, ov)
    names = ['M%d.omega' % c for c in range(n_cells)] + ['M%d.e_fd' % c for c in range(0, n_cells, 7)] + \
        ['bus:A3.v_re', 'bus:X5.v_im']
    rec = [(0, c, 0) for c in range(n_cells)] + [(0, c, 18) for c in range(0, n_cells, 7)] + \
        [(1, sim.bus_idxs['A3'], 0), (2, sim.bus_idxs['X5'], 0)]
    sim.set_record_channels(names)
    return sim, rec


def test_tiled_grid_host_pipeline_against_oracle():
    """Config-5 style tiled grid (here 20 cells = 100 buses / 20 generators: the wide-stepper regime, > 16 generators):
    prologue + planning + host build of the device code vs the C oracle on the same inputs."""
    from oracle import adapter
    sim, rec = _tiled_case(20, 3, 40, 'rk4', 'cpu')
    res, info = helpers.run_emulated(sim)
    ref = adapter.run_oracle_for(sim, info['init'], rec_channels=rec, full=False)
    assert np.abs(res - ref['rec'].transpose(2, 1, 0)).max() < 1e-9


def _init_fingerprint(init):
    parts = [init.Z0.detach().reshape(-1), init.Ydiag.detach().reshape(-1), init.bus_v.detach().reshape(-1)]
    flat = [torch.view_as_real(p.resolve_conj()).reshape(-1) if p.is_complex() else p for p in parts]
    for x0, sl in zip(init.x0, init.slots):
        flat += [v.detach().reshape(-1).to(torch.float64) for v in x0]
        flat += [sl[k].detach().reshape(-1).to(torch.float64) for k in sorted(sl)]
    return [f.clone() for f in flat]


def _same(a, b):
    return len(a) == len(b) and all(x.shape == y.shape and torch.equal(x, y) for x, y in zip(a, b))


def test_prologue_memoisation_is_invalidated_by_every_kind_of_change():
    """initialize() memoises the load flow, inv(Y_dyn) and every derived model constant on identity + version of their
    input tensors (prologue.py).  A cached run must be bit-identical to an uncached one, and every way of changing a
    model attribute -- rebinding a generator parameter, writing into a load tensor in place, replacing a controller
    parameter, changing a line -- must reach the result exactly as in a simulation built from scratch."""
    P = 5
    rng = np.random.RandomState(3)
    h_new = 6.5 * rng.uniform(0.8, 1.2, P)

    def fresh(mods):
        sim = Pss(parallel_sims=P, sim_time=0.1, time_step=0.005, solver='heun', grid_data=helpers.load_grid('k2a'),
                  verbose=False, device='cpu')
        for m in mods:
            m(sim)
        sim.reset()
        return sim

    def rebind_h(sim):
        sim.get_generator_by_name('G2').h = torch.tensor(h_new.reshape(P, 1), dtype=torch.complex128)

    def load_in_place(sim):
        load = next(m for b in sim.busses for m in b.models if type(m).__name__ == 'Load')
        load.p_soll_mw.mul_(1.03)                                   # same object, version counter bumped

    def avr_gain(sim):
        sim.get_generator_by_name('G3').exciter.pt1_lim.gain_pt1 = torch.full((P, 1), 150.0, dtype=torch.float64)

    def line_x(sim):
        sim.lines[2].x = sim.lines[2].x * 1.1

    sim = fresh([])
    base = _init_fingerprint(sim.initialize())
    sim.reset()
    assert _same(_init_fingerprint(sim.initialize()), base)          # pure cache hit
    applied = []
    for mod in (rebind_h, load_in_place, avr_gain, line_x):
        applied.append(mod)
        mod(sim)
        sim.reset()
        got = _init_fingerprint(sim.initialize())
        want = _init_fingerprint(fresh(applied).initialize())
        assert _same(got, want), mod.__name__
        assert not _same(got, base)
        sim.reset()
        assert _same(_init_fingerprint(sim.initialize()), got)      # and the new state is cached consistently


def test_network_reduce_torch_matches_device_code():
    """The differentiable torch twin of the fault re-reduction (planning.network_reduce_torch, used when a gradient flows
    into the network) equals the host build of the device code (csrc/dps_reduce.cuh) on a per-sim network with single and
    double faults."""
    import ctypes
    rng = np.random.RandomState(0)
    B, n = 3, 7
    y = rng.randn(B, n, n) + 1j * rng.randn(B, n, n)
    y = y + np.transpose(y, (0, 2, 1)) + (8 + 4j) * np.eye(n)
    Z0 = torch.linalg.inv(torch.from_numpy(y))
    Yd = torch.diagonal(torch.from_numpy(y), dim1=1, dim2=2)
    obs_bus, gen_bus, cfg_sets, m_max = [1, 4, 6, 2], [1, 4, 6], [(), (3,), (0, 5)], 2
    mine = planning.network_reduce_torch(Z0, Yd, obs_bus, gen_bus, cfg_sets, m_max).numpy()
    lib = helpers.build_emul()
    z_re, z_im = (a.numpy() for a in planning.split_complex_batch_last(Z0))
    yd_re, yd_im = (a.numpy() for a in planning.split_complex_batch_last(Yd))
    ref = np.zeros((len(cfg_sets), len(obs_bus) + m_max, len(gen_bus) + m_max, 2, B))
    ci = ctypes.c_int
    ptr, flat = [0], []
    for s in cfg_sets:
        flat += list(s)
        ptr.append(len(flat))
    p_ = helpers._p
    lib.emul_network_reduce(ctypes.c_longlong(B), n, p_(z_re), p_(z_im), p_(yd_re), p_(yd_im), len(obs_bus),
                            (ci * len(obs_bus))(*obs_bus), len(gen_bus), (ci * len(gen_bus))(*gen_bus), len(cfg_sets),
                            (ci * len(ptr))(*ptr), (ci * len(flat))(*flat), ctypes.c_double(planning.FAULT_ADMITTANCE), p_(ref))
    no, ng = len(obs_bus), len(gen_bus)
    assert np.abs(mine[:, :no, :ng] - ref[:, :no, :ng]).max() < 1e-12


@pytest.mark.parametrize('grid,names', [
    ('k2a', [('G1', 'h'), ('G2', 'd'), ('G3', 'x_d_t'), ('G4', 'x_q')]),
    ('k2a', [('G1', 'x_d_st'), ('G2', 's_n')]),
    ('ibb_wc', [('G1', 'x_d'), ('G1', 'exciter.pt1_lim.gain_pt1')]),
    ('k2a', [('G2', 'governor.pt1_lim.gain_pt1'), ('G3', 'stabilizer.washout.k_w'), ('G1', 'exciter.lead_lag.t_1'),
             ('G4', 'stabilizer.lead_lag2.t_2'), ('G4', 'governor.lead_lag.t_2')]),
    ('ieee9', [('G2', 'h'), ('G3', 'x_q_st'), ('G1', 't_d0_t')]),
])
def test_device_prologue_classification_matches_broadcasting(grid, names):
    """The kernel prologue (prologue_device.py) decides which initial states and constant slots differ per trajectory from a
    static dependency table; the torch prologue gets the same answer from broadcasting.  Host logic: checked here without a GPU."""
    from diffpssi_b200 import prologue_device as pd
    P = 5
    sim = Pss(parallel_sims=P, sim_time=0.1, time_step=0.005, solver='heun', grid_data=helpers.load_grid(grid), verbose=False,
              device='cpu')
    rng = np.random.RandomState(3)
    ov = {}
    for gname, attr in names:
        obj, name = helpers.resolve_attr(sim, gname, attr)
        old = getattr(obj, name)
        base = float(old.reshape(-1)[0].real) if torch.is_tensor(old) else float(old)
        ov[(gname, attr)] = base * rng.uniform(0.95, 1.05, P)
    helpers.apply_overrides(sim, ov)
    init = sim.initialize()                                                      # torch path (no CUDA device here)
    col = prologue._Cols(P, 'cpu')
    gens = prologue.generators_of(sim)
    vary = tuple(a is not None and col(a, torch.float64).shape[0] > 1 for _, g in gens for a in pd._raw_attrs(g))
    base_vary = init.bus_v.shape[0] > 1 or init.s_calc.shape[0] > 1
    cls = pd._classification(vary, base_vary, init.ctrl)
    for gi, (rows, xv, slot_names, sv) in enumerate(cls):
        assert len(rows) == len(init.x0[gi])
        for k, t in zip(rows, init.x0[gi]):
            if k in (0, 10, 11, 12):
                continue                         # identically zero rows: their batch shape carries no information
            assert (t.shape[0] > 1) == xv[k], (gi, k)
        assert set(slot_names) == set(init.slots[gi])
        for name in slot_names:
            assert (init.slots[gi][name].shape[0] > 1) == sv[name], (gi, name)
