"""Builder-authored grid dicts in the reference's dict schema (simulator.py:128-209).

The reference ships IBB, IEEE 9-bus and Kundur two-area grids only (examples/models/*).  BASELINE.json's headline
configuration is an IEEE 39-bus / 10-generator system with SEXS + TGOV1 + STAB1 on every machine, which does not exist
in the reference -- it is authored here with the same keys so it loads in the reference too (the CPU baselines in
bench.py run the oracle on exactly this dict).

Network data: the New-England 39-bus test system (branch r/x/b and loads on a 100 MVA base, as tabulated in the
literature; transformer taps are taken as nominal because the reference's Transformer model has none).  Machine data:
round-rotor parameters on each machine's own MVA base in the style of the reference's K2A example.

The reference initialises a machine from the *net* injection of its bus (simulator.py:443-445 hands S_calc[:, i] to
every model on bus i), so a load sharing a bus with a generator would be counted twice (once in p_m, once as a
constant admittance) and the run would not start in equilibrium.  The two such loads of the original data set are
therefore netted into their machines: the 9.2 MW at the slack bus 31 is dropped (the slack picks it up) and the
1104 MW at bus 39 turns the 1000 MW external-system equivalent G10 into a net -104 MW machine -- the same device the
reference's own IBB example uses for its infinite bus (examples/models/ibb_model/ibb_model.py: P = -1998).
"""


def ieee39():
    bus = ['B%d' % i for i in range(1, 40)]
    # (from, to, r, x, b) lines; b = 0 rows with a bus >= 30 or the 11-12-13 / 19-20 links are transformers
    branches = [
        (1, 2, .0035, .0411, .6987), (1, 39, .0010, .0250, .7500), (2, 3, .0013, .0151, .2572), (2, 25, .0070, .0086, .1460),
        (3, 4, .0013, .0213, .2214), (3, 18, .0011, .0133, .2138), (4, 5, .0008, .0128, .1342), (4, 14, .0008, .0129, .1382),
        (5, 6, .0002, .0026, .0434), (5, 8, .0008, .0112, .1476), (6, 7, .0006, .0092, .1130), (6, 11, .0007, .0082, .1389),
        (7, 8, .0004, .0046, .0780), (8, 9, .0023, .0363, .3804), (9, 39, .0010, .0250, 1.2000), (10, 11, .0004, .0043, .0729),
        (10, 13, .0004, .0043, .0729), (13, 14, .0009, .0101, .1723), (14, 15, .0018, .0217, .3660), (15, 16, .0009, .0094, .1710),
        (16, 17, .0007, .0089, .1342), (16, 19, .0016, .0195, .3040), (16, 21, .0008, .0135, .2548), (16, 24, .0003, .0059, .0680),
        (17, 18, .0007, .0082, .1319), (17, 27, .0013, .0173, .3216), (21, 22, .0008, .0140, .2565), (22, 23, .0006, .0096, .1846),
        (23, 24, .0022, .0350, .3610), (25, 26, .0032, .0323, .5130), (26, 27, .0014, .0147, .2396), (26, 28, .0043, .0474, .7802),
        (26, 29, .0057, .0625, 1.0290), (28, 29, .0014, .0151, .2490),
    ]
    trafos = [
        (2, 30, .0000, .0181), (6, 31, .0000, .0250), (10, 32, .0000, .0200), (12, 11, .0016, .0435), (12, 13, .0016, .0435),
        (19, 20, .0007, .0138), (19, 33, .0007, .0142), (20, 34, .0009, .0180), (22, 35, .0000, .0143), (23, 36, .0005, .0272),
        (25, 37, .0006, .0232), (29, 38, .0008, .0156),
    ]
    loads = [
        (1, 97.6, 44.2), (3, 322.0, 2.4), (4, 500.0, 184.0), (7, 233.8, 84.0), (8, 522.0, 176.6), (9, 6.5, -66.6),
        (12, 8.53, 88.0), (15, 320.0, 153.0), (16, 329.0, 32.3), (18, 158.0, 30.0), (20, 680.0, 103.0), (21, 274.0, 115.0),
        (23, 247.5, 84.6), (24, 308.6, -92.2), (25, 224.0, 47.2), (26, 139.0, 17.0), (27, 281.0, 75.5), (28, 206.0, 27.6),
        (29, 283.5, 26.9),
    ]
    # (bus, S_n, P, V, H, X_d, X_q, X_d', X_q', X'')
    gens = [
        (30, 1000, 250.0, 1.0475, 4.20, 1.80, 1.70, 0.30, 0.55, 0.25),
        (31, 700, 520.8, 0.9820, 4.33, 1.85, 1.75, 0.32, 0.57, 0.24),
        (32, 800, 650.0, 0.9831, 4.475, 1.75, 1.68, 0.29, 0.52, 0.23),
        (33, 800, 632.0, 0.9972, 3.575, 1.82, 1.74, 0.31, 0.56, 0.25),
        (34, 600, 508.0, 1.0123, 4.33, 1.90, 1.80, 0.33, 0.60, 0.26),
        (35, 800, 650.0, 1.0493, 4.35, 1.78, 1.70, 0.30, 0.54, 0.24),
        (36, 700, 560.0, 1.0635, 3.77, 1.84, 1.76, 0.31, 0.55, 0.25),
        (37, 700, 540.0, 1.0278, 3.47, 1.80, 1.72, 0.30, 0.55, 0.25),
        (38, 1000, 830.0, 1.0265, 3.45, 1.88, 1.79, 0.32, 0.58, 0.26),
        (39, 10000, -104.0, 1.0300, 5.00, 1.80, 1.70, 0.30, 0.55, 0.25),
    ]
    grid = {
        'base_mva': 100, 'f': 60, 'slack_bus': 'B31', 'base_voltage': 345,
        'busses': [['name', 'V_n']] + [[b, 345] for b in bus],
        'lines': [['name', 'from_bus', 'to_bus', 'length', 'S_n', 'V_n', 'unit', 'R', 'X', 'B']] +
                 [['L%d-%d' % (f, t), 'B%d' % f, 'B%d' % t, 1, 100, 345, 'p.u.', r, x, b] for f, t, r, x, b in branches],
        'transformers': [['name', 'from_bus', 'to_bus', 'S_n', 'V_n_from', 'V_n_to', 'R', 'X']] +
                        [['T%d-%d' % (f, t), 'B%d' % f, 'B%d' % t, 100, 345, 345, r, x] for f, t, r, x in trafos],
        'loads': [['name', 'bus', 'P', 'Q', 'model']] + [['LD%d' % b, 'B%d' % b, p, q, 'Z'] for b, p, q in loads],
        'generators': {'GEN': [['name', 'bus', 'S_n', 'V_n', 'P', 'V', 'H', 'D', 'X_d', 'X_q', 'X_d_t', 'X_q_t', 'X_d_st',
                                'X_q_st', 'T_d0_t', 'T_q0_t', 'T_d0_st', 'T_q0_st']] +
                               [['G%d' % (i + 1), 'B%d' % b, sn, 345, p, v, h, 0, xd, xq, xdt, xqt, xst, xst, 8.0, 0.4, 0.03, 0.05]
                                for i, (b, sn, p, v, h, xd, xq, xdt, xqt, xst) in enumerate(gens)]},
        'gov': {'TGOV1': [['name', 'gen', 'R', 'D_t', 'V_min', 'V_max', 'T_1', 'T_2', 'T_3']] +
                          # V_min = -1 everywhere: the net-import equivalent G10 runs at p_m < 0, and identical
                          # controller settings on all machines let the stepper use its constant-bank path
                          [['GOV%d' % (i + 1), 'G%d' % (i + 1), 0.05, 0.02, -1, 1, 0.5, 1, 2] for i in range(10)]},
        'avr': {'SEXS': [['name', 'gen', 'K', 'T_a', 'T_b', 'T_e', 'E_min', 'E_max']] +
                         [['AVR%d' % (i + 1), 'G%d' % (i + 1), 100, 2.0, 10.0, 0.1, -5, 5] for i in range(10)]},
        'pss': {'STAB1': [['name', 'gen', 'K', 'T', 'T_1', 'T_2', 'T_3', 'T_4', 'H_lim']] +
                          [['PSS%d' % (i + 1), 'G%d' % (i + 1), 50, 10.0, 0.5, 0.5, 0.05, 0.05, 0.03] for i in range(10)]},
    }
    return grid


GENERATOR_NOMINALS = {
    # (H, X_d') per generator, for the "U(0.5,2) x nominal" synthetic parameter batches of SURVEY.md section 8(d)
    'ieee39': {'G%d' % (i + 1): v for i, v in enumerate([
        (4.20, 0.30), (4.33, 0.32), (4.475, 0.29), (3.575, 0.31), (4.33, 0.33), (4.35, 0.30), (3.77, 0.31), (3.47, 0.30),
        (3.45, 0.32), (5.00, 0.30)])},
    'k2a': {'G1': (6.5, 0.3), 'G2': (6.5, 0.3), 'G3': (6.175, 0.3), 'G4': (6.175, 0.3)},
}


def tiled(n_cells=100, ring_stride=10):
    """Synthetic tiled grid for BASELINE.json config 5: ``n_cells`` identical 5-bus / 1-generator cells (500 buses and
    100 generators for the default) with all three controllers on every machine, cells tied into a ring plus a second
    ring with stride ``ring_stride`` (a torus-like mesh).  The network is shared by the batch; H, D, X_d' may vary.

    Cell c: G<c> (generator, 20 kV) --T-- H<c> (230 kV) -- A<c>, B<c> (loads) -- X<c> (tie bus);
    ties: X<c> -- H<c+1>  and  X<c> -- X<c+ring_stride>  (indices modulo n_cells).
    """
    n = n_cells
    busses, lines, trafos, loads, gens = [], [], [], [], []
    for c in range(n):
        busses += [['G%d' % c, 20], ['H%d' % c, 230], ['A%d' % c, 230], ['B%d' % c, 230], ['X%d' % c, 230]]
        trafos.append(['T%d' % c, 'G%d' % c, 'H%d' % c, 900, 20, 230, 0, 0.15])
        for nm, f, t, length in (('LHA', 'H', 'A', 20), ('LHB', 'H', 'B', 30), ('LAX', 'A', 'X', 25), ('LBX', 'B', 'X', 15)):
            lines.append(['%s%d' % (nm, c), '%s%d' % (f, c), '%s%d' % (t, c), length, 100, 230, 'p.u.', 1e-4, 1e-3, 1.75e-3])
        lines.append(['TIE%d' % c, 'X%d' % c, 'H%d' % ((c + 1) % n), 40, 100, 230, 'p.u.', 1e-4, 1e-3, 1.75e-3])
        if n > ring_stride:
            lines.append(['MESH%d' % c, 'X%d' % c, 'X%d' % ((c + ring_stride) % n), 60, 100, 230, 'p.u.', 1e-4, 1e-3, 1.75e-3])
        # alternate heavier / lighter cells so power actually flows over the ties
        heavy = (c % 2 == 0)
        loads.append(['LA%d' % c, 'A%d' % c, 375.0 if heavy else 295.0, 60.0, 'Z'])
        loads.append(['LB%d' % c, 'B%d' % c, 335.0 if heavy else 315.0, 50.0, 'Z'])
        gens.append(['M%d' % c, 'G%d' % c, 900, 20, 690.0 if heavy else 640.0, 1.02 if heavy else 1.01, 6.5 if heavy else 6.175, 0,
                     1.8, 1.7, 0.3, 0.55, 0.25, 0.25, 8.0, 0.4, 0.03, 0.05])
    return {
        'base_mva': 900, 'f': 50, 'slack_bus': 'G0', 'base_voltage': 230,
        'busses': [['name', 'V_n']] + busses,
        'lines': [['name', 'from_bus', 'to_bus', 'length', 'S_n', 'V_n', 'unit', 'R', 'X', 'B']] + lines,
        'transformers': [['name', 'from_bus', 'to_bus', 'S_n', 'V_n_from', 'V_n_to', 'R', 'X']] + trafos,
        'loads': [['name', 'bus', 'P', 'Q', 'model']] + loads,
        'generators': {'GEN': [['name', 'bus', 'S_n', 'V_n', 'P', 'V', 'H', 'D', 'X_d', 'X_q', 'X_d_t', 'X_q_t', 'X_d_st',
                                'X_q_st', 'T_d0_t', 'T_q0_t', 'T_d0_st', 'T_q0_st']] + gens},
        'gov': {'TGOV1': [['name', 'gen', 'R', 'D_t', 'V_min', 'V_max', 'T_1', 'T_2', 'T_3']] +
                          [['GOV%d' % c, 'M%d' % c, 0.05, 0.02, 0, 1, 0.5, 1, 2] for c in range(n)]},
        'avr': {'SEXS': [['name', 'gen', 'K', 'T_a', 'T_b', 'T_e', 'E_min', 'E_max']] +
                         [['AVR%d' % c, 'M%d' % c, 100, 2.0, 10.0, 0.1, -3, 3] for c in range(n)]},
        'pss': {'STAB1': [['name', 'gen', 'K', 'T', 'T_1', 'T_2', 'T_3', 'T_4', 'H_lim']] +
                          [['PSS%d' % c, 'M%d' % c, 50, 10.0, 0.5, 0.5, 0.05, 0.05, 0.03] for c in range(n)]},
    }
