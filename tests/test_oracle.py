"""Pins the C oracle (oracle/dps_oracle.c) against the reference.

(a) the reference's own golden arrays (its tests/integration/data/*.npy, numpy backend, tolerance 1e-6 there);
(b) live outputs of the reference's torch backend for every state, algebraic and bus voltage, all three solvers,
    per-sim parameter batches, a per-sim network and multi-event schedules (tests/golden/*.npz).
The oracle starts from the reference's own post-initialize() snapshot stored in the fixtures, so this file tests the
restatement of the *time loop* only and involves no product code.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DT = 0.005


def load_grid(name):
    with open(os.path.join(GOLD, 'grids', name + '.json')) as f:
        return json.load(f)


def run_case(case, grid_name):
    gold = np.load(os.path.join(GOLD, case + '.npz'))
    grid = load_grid(grid_name)
    P = gold['traj'].shape[0]
    inp = oracle.inputs_from_golden(grid, gold, P)
    t = gold['time']
    out = oracle.run(inp['n_bus'], inp['gen_bus'], inp['ctrl'], str(gold['solver']), DT, inp['fn'], t, inp['events'],
                     inp['raw'], inp['x0'], inp['z0'])
    chans = oracle.golden_channel_table(gold, inp['gen_names'], inp['bus_names'], inp['ctrl'])
    steps = gold['steps']
    mine = np.zeros_like(gold['traj'])
    for c, (kind, idx, field) in enumerate(chans):
        if kind == 0:
            mine[:, :, c] = out['full'][steps, idx, field, :].T
        else:
            mine[:, :, c] = out['busv'][steps, idx, kind - 1, :].T
    return gold, mine, out, inp


EXAMPLES = [('ibb', 'ibb_model'), ('ibb_trans', 'ibb_transformer'), ('ibb_wc', 'ibb_with_controllers'),
            ('ieee9', 'ieee_9bus'), ('k2a', 'k2a')]


@pytest.mark.parametrize('grid_name,ref_file', EXAMPLES)
@pytest.mark.parametrize('solver', ['heun', 'euler', 'rk4'])
def test_oracle_matches_reference_torch_path(grid_name, ref_file, solver):
    gold, mine, out, inp = run_case('%s_%s' % (grid_name, solver), grid_name)
    err = np.abs(mine - gold['traj']).max()
    assert err < 1e-9, err            # north-star tolerance: 1e-9 p.u. absolute over the full simulation


@pytest.mark.parametrize('grid_name,ref_file', EXAMPLES)
def test_oracle_matches_reference_golden_arrays(grid_name, ref_file):
    """The reference's own fixtures (made by its numpy backend); tolerance as tests/integration/test_example_sims.py:18-19."""
    gold, mine, out, inp = run_case('%s_heun' % grid_name, grid_name)
    ref = np.load(os.path.join(GOLD, 'ref_%s.npy' % ref_file))
    names = [str(s) for s in gold['names']]
    chans = oracle.golden_channel_table(gold, inp['gen_names'], inp['bus_names'], inp['ctrl'])
    cols = []
    for c in [str(s) for s in gold['example_channels']]:
        kind, idx, field = chans[names.index(c)]
        cols.append(out['full'][:, idx, field, 0] if kind == 0 else out['busv'][:, idx, kind - 1, 0])
    mine_rec = np.stack(cols, axis=1)
    assert np.allclose(mine_rec, ref, atol=1e-6, rtol=1e-6)
    assert np.abs(mine_rec - gold['example_rec']).max() < 1e-9


@pytest.mark.parametrize('case,grid_name,tol', [
    ('k2a_batch_heun', 'k2a', 1e-9),
    ('ieee9_persimz_heun', 'ieee9', 1e-9),
    ('ieee9_two_events_heun', 'ieee9', 1e-9),
    ('k2a_two_events_rk4', 'k2a', 1e-9),
    # the configurations bench.py and BASELINE.json name (builder-authored grids in the reference's schema, run through
    # the unmodified reference by oracle/gen_golden.py: headline_cases)
    ('ieee39_heun', 'ieee39', 1e-9),                 # config 3 grid, nominal, full 2000 Heun steps
    ('ieee39_batch_heun', 'ieee39', 1e-9),           # rows 0..3 of the bench's own U(0.5,2) parameter draw
    ('tiled20_rk4', 'tiled20', 1e-9),                # config 5 family: 100 buses / 20 generators, RK4
    ('tiled100_rk4', 'tiled100', 1e-9),              # config 5 grid itself: 500 buses / 100 generators, RK4
])
def test_oracle_batches_and_events(case, grid_name, tol):
    gold, mine, out, inp = run_case(case, grid_name)
    err = np.abs(mine - gold['traj']).max()
    assert err < tol, err


def test_oracle_ibb_identification_range():
    """0.5x..2x parameter range of the reference's ibb_opt.py: contains pole-slipping rows (|delta| reaches 1e2 rad)
    whose round-off amplification exceeds 1e-9 even between the reference's own two backends (SURVEY.md hard part 1),
    so parity is asserted per row: stable rows to 1e-9, all rows to 1e-6 relative."""
    gold, mine, out, inp = run_case('ibb_batch_heun', 'ibb')
    names = [str(s) for s in gold['names']]
    d = gold['traj'][:, :, names.index('Gen 1.delta')]
    stable = np.abs(d).max(axis=1) < 10.0
    assert stable.sum() >= 4
    err = np.abs(mine - gold['traj']).max(axis=(1, 2))
    assert (err[stable] < 1e-9).all(), err
    scale = np.abs(gold['traj']).max(axis=(1, 2))
    assert (err / scale < 1e-6).all(), err / scale
