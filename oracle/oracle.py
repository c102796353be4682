"""TEST INFRASTRUCTURE -- numpy/ctypes front end of the C oracle (oracle/dps_oracle.c).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs import this.
It never touches the CUDA engine and the product package never imports it.
"""
import ctypes
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'dps_oracle.c')
LIB = os.path.join(HERE, 'libdps_oracle.so')

N_STATES, N_FIELDS, N_RAW = 13, 20, 39
RAW = ('h', 'd', 'x_d', 'x_q', 'x_d_t', 'x_q_t', 'x_d_st', 'x_q_st', 't_d0_t', 't_q0_t', 't_d0_st', 't_q0_st',
       's_n', 's_n_sys', 'p_m0', 'e_fd0',
       'a_t_a', 'a_t_b', 'a_t_e', 'a_k', 'a_e_min', 'a_e_max', 'a_v_set', 'a_bias',
       'g_t_1', 'g_gain', 'g_v_min', 'g_v_max', 'g_t_2', 'g_t_3', 'g_d_t', 'g_p_ref',
       's_k_w', 's_t_w', 's_t_1', 's_t_2', 's_t_3', 's_t_4', 's_h_lim')
RAW_INDEX = {n: i for i, n in enumerate(RAW)}
assert len(RAW) == N_RAW
SOLVER_ID = {'euler': 0, 'heun': 1, 'rk4': 2}
STATE_NAMES = ['omega', 'delta', 'e_q_t', 'e_d_t', 'e_q_st', 'e_d_st', 'avr_ll', 'avr_pt1', 'gov_pt1', 'gov_ll',
               'pss_w', 'pss_ll1', 'pss_ll2']
FIELD_NAMES = STATE_NAMES + ['p_e', 'i_d', 'i_q', 'v_re', 'v_im', 'e_fd', 'p_m']
CTRL_AVR, CTRL_GOV, CTRL_PSS = 1, 2, 4

# generator-dict key -> raw slot; controller override attribute paths -> raw slot
GEN_KEYS = {'H': 'h', 'D': 'd', 'X_d': 'x_d', 'X_q': 'x_q', 'X_d_t': 'x_d_t', 'X_q_t': 'x_q_t', 'X_d_st': 'x_d_st',
            'X_q_st': 'x_q_st', 'T_d0_t': 't_d0_t', 'T_q0_t': 't_q0_t', 'T_d0_st': 't_d0_st', 'T_q0_st': 't_q0_st',
            'S_n': 's_n'}
ATTR_TO_RAW = {
    'h': 'h', 'd': 'd', 'x_d': 'x_d', 'x_q': 'x_q', 'x_d_t': 'x_d_t', 'x_q_t': 'x_q_t', 'x_d_st': 'x_d_st',
    'x_q_st': 'x_q_st', 't_d0_t': 't_d0_t', 't_q0_t': 't_q0_t', 't_d0_st': 't_d0_st', 't_q0_st': 't_q0_st',
    's_n': 's_n',
    'exciter.lead_lag.t_1': 'a_t_a', 'exciter.lead_lag.t_2': 'a_t_b', 'exciter.pt1_lim.t_pt1': 'a_t_e',
    'exciter.pt1_lim.gain_pt1': 'a_k', 'exciter.pt1_lim.lim_min': 'a_e_min', 'exciter.pt1_lim.lim_max': 'a_e_max',
    'governor.pt1_lim.t_pt1': 'g_t_1', 'governor.pt1_lim.gain_pt1': 'g_gain', 'governor.pt1_lim.lim_min': 'g_v_min',
    'governor.pt1_lim.lim_max': 'g_v_max', 'governor.lead_lag.t_1': 'g_t_2', 'governor.lead_lag.t_2': 'g_t_3',
    'governor.d_t': 'g_d_t',
    'stabilizer.washout.k_w': 's_k_w', 'stabilizer.washout.t_w': 's_t_w', 'stabilizer.lead_lag1.t_1': 's_t_1',
    'stabilizer.lead_lag1.t_2': 's_t_3', 'stabilizer.lead_lag2.t_1': 's_t_2', 'stabilizer.lead_lag2.t_2': 's_t_4',
    'stabilizer.limiter.limit': 's_h_lim',
}


class Topo(ctypes.Structure):
    _fields_ = [('n_bus', ctypes.c_int), ('n_gen', ctypes.c_int), ('solver', ctypes.c_int), ('n_steps', ctypes.c_int),
                ('dt', ctypes.c_double), ('fn', ctypes.c_double),
                ('gen_bus', ctypes.POINTER(ctypes.c_int)), ('ctrl', ctypes.POINTER(ctypes.c_int)),
                ('n_events', ctypes.c_int), ('ev_bus', ctypes.POINTER(ctypes.c_int)),
                ('ev_active', ctypes.POINTER(ctypes.c_ubyte)),
                ('n_rec', ctypes.c_int), ('rec', ctypes.POINTER(ctypes.c_int))]


def build(force=False):
    """Compile the C oracle with gcc (seconds).  Returns the path of the shared library."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(['gcc', '-O2', '-fopenmp', '-fcx-limited-range', '-shared', '-fPIC', '-o', LIB, SRC, '-lm'])
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.dps_oracle_run.restype = ctypes.c_int
    return _lib


def event_flags(time_grid, events):
    """[T, n_events] uint8: ``start < t <= end`` on the reference's float time grid (static_models.py:46)."""
    act = np.zeros((len(time_grid), max(len(events), 1)), dtype=np.uint8)
    for e, (t0, t1, _b) in enumerate(events):
        act[:, e] = [(t0 < t <= t1) for t in time_grid]
    return act


def run(n_bus, gen_bus, ctrl, solver, dt, fn, time_grid, events, raw, x0, z0, rec=(), full=True, busv=True,
        threads=None):
    """events: [(t0, t1, bus_index)].  raw [G,N_RAW,P], x0 [G,13,P], z0 complex [Bz,n,n].
    Returns dict(rec [n_rec,T,P], full [T,G,20,P], busv [T,n,2,P])."""
    lib = _load()
    G, T = len(gen_bus), len(time_grid)
    P = x0.shape[2]
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    z0 = np.ascontiguousarray(z0, dtype=np.complex128)
    assert raw.shape == (G, N_RAW, P) and x0.shape == (G, N_STATES, P) and z0.shape[1:] == (n_bus, n_bus)
    Bz = z0.shape[0]
    act = np.ascontiguousarray(event_flags(time_grid, events))
    ci = ctypes.c_int
    a_gb = (ci * G)(*gen_bus)
    a_ct = (ci * G)(*ctrl)
    a_eb = (ci * max(len(events), 1))(*[e[2] for e in events] or [0])
    rec = list(rec)
    a_rec = (ci * max(3 * len(rec), 1))(*[v for c in rec for v in c] or [0])
    topo = Topo(n_bus, G, SOLVER_ID[solver], T, dt, fn, a_gb, a_ct, len(events), a_eb,
                act.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)), len(rec), a_rec)
    out = {}
    o_rec = np.zeros((len(rec), T, P)) if rec else None
    o_full = np.zeros((T, G, N_FIELDS, P)) if full else None
    o_bus = np.zeros((T, n_bus, 2, P)) if busv else None

    def ptr(a):
        return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None

    rc = lib.dps_oracle_run(ctypes.byref(topo), ctypes.c_long(P), ptr(raw), ptr(x0), ptr(z0.view(np.float64)),
                            ctypes.c_long(Bz), ptr(o_rec), ptr(o_full), ptr(o_bus),
                            ctypes.c_int(threads or os.cpu_count() or 1))
    if rc != 0:
        raise RuntimeError('oracle: singular matrix during fault switching')
    out['rec'], out['full'], out['busv'] = o_rec, o_full, o_bus
    return out


# ----------------------------------------------------------------------------------------------------------------------
# building oracle inputs straight from a grid dict + a golden snapshot (no product code involved)
# ----------------------------------------------------------------------------------------------------------------------
def _table(tab):
    keys = tab[0]
    return [dict(zip(keys, row)) for row in tab[1:]]


def inputs_from_golden(grid, gold, P):
    """Raw parameters / x0 / Z0 for the oracle from a grid dict and a golden .npz made by oracle/gen_golden.py
    (which stores the reference's own state after initialize()).  Per-sim overrides stored in the npz are applied."""
    bus_names = [b['name'] for b in _table(grid['busses'])]
    gens = _table(grid['generators']['GEN'])
    order = sorted(range(len(gens)), key=lambda i: (bus_names.index(gens[i]['bus']), i))   # bus order, then model order
    gens = [gens[i] for i in order]
    avr = {d['gen']: d for d in _table(grid['avr']['SEXS'])} if isinstance(grid.get('avr'), dict) else {}
    gov = {d['gen']: d for d in _table(grid['gov']['TGOV1'])} if isinstance(grid.get('gov'), dict) else {}
    pss = {d['gen']: d for d in _table(grid['pss']['STAB1'])} if isinstance(grid.get('pss'), dict) else {}
    G = len(gens)
    raw = np.zeros((G, N_RAW, P))
    x0 = np.zeros((G, N_STATES, P))
    ctrl, gen_bus = [], []

    def put(g, name, val):
        raw[g, RAW_INDEX[name], :] = val

    for g, d in enumerate(gens):
        name = d['name']
        gen_bus.append(bus_names.index(d['bus']))
        for k, rn in GEN_KEYS.items():
            put(g, rn, d[k])
        put(g, 's_n_sys', grid['base_mva'])
        put(g, 'p_m0', gold['init/gen:%s.p_m' % name])
        put(g, 'e_fd0', gold['init/gen:%s.e_fd' % name])
        flags = 0
        xs = gold['init/gen:%s.x0' % name]                      # [P, s] packed: machine | avr | gov | pss
        x0[g, :6, :] = xs[:, :6].T
        off = 6
        if name in avr:
            a = avr[name]
            flags |= CTRL_AVR
            for rn, k in (('a_t_a', 'T_a'), ('a_t_b', 'T_b'), ('a_t_e', 'T_e'), ('a_k', 'K'), ('a_e_min', 'E_min'),
                          ('a_e_max', 'E_max')):
                put(g, rn, a[k])
            put(g, 'a_v_set', gold['init/gen:%s.avr_vset' % name])
            put(g, 'a_bias', gold['init/gen:%s.avr_bias' % name])
            x0[g, 6:8, :] = xs[:, off:off + 2].T
            off += 2
        if name in gov:
            v = gov[name]
            flags |= CTRL_GOV
            for rn, k in (('g_t_1', 'T_1'), ('g_v_min', 'V_min'), ('g_v_max', 'V_max'), ('g_t_2', 'T_2'),
                          ('g_t_3', 'T_3'), ('g_d_t', 'D_t')):
                put(g, rn, v[k])
            put(g, 'g_gain', 1 / v['R'])                         # governors.py:71
            put(g, 'g_p_ref', gold['init/gen:%s.gov_pref' % name])
            x0[g, 8:10, :] = xs[:, off:off + 2].T
            off += 2
        if name in pss:
            s = pss[name]
            flags |= CTRL_PSS
            for rn, k in (('s_k_w', 'K'), ('s_t_w', 'T'), ('s_t_1', 'T_1'), ('s_t_2', 'T_2'), ('s_t_3', 'T_3'),
                          ('s_t_4', 'T_4'), ('s_h_lim', 'H_lim')):
                put(g, rn, s[k])
            x0[g, 10:13, :] = xs[:, off:off + 3].T
        ctrl.append(flags)
    names = [d['name'] for d in gens]
    for key in gold.files:
        if key.startswith('override/'):
            _kind, gname, attr = key[len('override/'):].split('|')
            raw[names.index(gname), RAW_INDEX[ATTR_TO_RAW[attr]], :] = gold[key]
    z0 = gold['init/Z0']
    events = [(t0, t1, bus_names.index(b)) for (t0, t1, b) in json.loads(str(gold['events']))]
    return dict(n_bus=len(bus_names), gen_bus=gen_bus, ctrl=ctrl, raw=raw, x0=x0, z0=z0, events=events,
                gen_names=names, bus_names=bus_names, fn=grid['f'])


def golden_channel_table(gold, gen_names, bus_names, ctrl):
    """Map every channel name of a golden trajectory to (kind, index, field) of the oracle/engine vocabulary."""
    out = []
    for n in [str(s) for s in gold['names']]:
        if n.startswith('bus:'):
            b, f = n[4:].rsplit('.', 1)
            out.append((1 if f == 'v_re' else 2, bus_names.index(b), 0))
        else:
            g, f = n.rsplit('.', 1)
            out.append((0, gen_names.index(g), FIELD_NAMES.index(f)))
    return out
