"""TEST INFRASTRUCTURE -- not product code.

Thin driver around the *unmodified* reference at /root/reference (read-only,
georgkordowich/diffpssi 0.0.3).  It exists only in this container: it is used by
``oracle/gen_golden.py`` to produce the fixtures under ``tests/golden/`` and by
optional cross-checks (tests marked ``needs_reference``).  Nothing in the
product package, ``bench.py`` or the ``-m gpu`` tests imports it at run time.

What it drives (reference file:line):
  * ``PowerSystemSimulation.__init__/create_grid``  simulator.py:59-209
  * ``initialize()``                                simulator.py:438-450
  * ``run()`` hot loop                              simulator.py:497-552
with the reference's *torch* backend (``DIFFPSSI_FORCE_SIM_BACKEND=torch`` must be
set before first import, backend.py:11-14) because parity is judged against the
torch fp64 path.
"""
import importlib.util
import os
import sys

REFERENCE_ROOT = '/root/reference'

# Channel vocabulary shared with diffpssi_b200.channels (kept in sync by tests).
STATE_NAMES = ['omega', 'delta', 'e_q_t', 'e_d_t', 'e_q_st', 'e_d_st']
AVR_NAMES = ['avr_ll', 'avr_pt1']
GOV_NAMES = ['gov_pt1', 'gov_ll']
PSS_NAMES = ['pss_w', 'pss_ll1', 'pss_ll2']
ALG_NAMES = ['p_e', 'i_d', 'i_q', 'v_re', 'v_im', 'e_fd', 'p_m']


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'src', 'diffpssi'))


def _import_reference(backend='torch'):
    """Import the reference simulator module with the requested backend.

    The backend is a module-level global fixed at first import (backend.py:8-14), so one
    process can only ever host one backend.
    """
    want = backend
    have = os.environ.get('DIFFPSSI_FORCE_SIM_BACKEND')
    if 'src.diffpssi.power_sim_lib.backend' in sys.modules:
        mod = sys.modules['src.diffpssi.power_sim_lib.backend']
        if mod.BACKEND != want:
            raise RuntimeError('reference already imported with backend %s' % mod.BACKEND)
    else:
        if have is not None and have != want:
            raise RuntimeError('DIFFPSSI_FORCE_SIM_BACKEND=%s conflicts with %s' % (have, want))
        os.environ['DIFFPSSI_FORCE_SIM_BACKEND'] = want
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
    import src.diffpssi.power_sim_lib.simulator as simulator  # noqa
    return simulator


def load_example_grid(example, module):
    """Load a grid dict from /root/reference/examples/models/<example>/<module>.py."""
    path = os.path.join(REFERENCE_ROOT, 'examples', 'models', example, module + '.py')
    spec = importlib.util.spec_from_file_location('_refgrid_' + module, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.load()


def resolve(sim, path):
    """('gen', 'G1', 'stabilizer.washout.k_w') -> (object, attribute)."""
    kind, name, attr = path
    obj = None
    for bus in sim.busses:
        for m in bus.models:
            if getattr(m, 'name', None) == name:
                obj = m
    if obj is None:
        raise KeyError(name)
    parts = attr.split('.')
    for p in parts[:-1]:
        obj = getattr(obj, p)
    return obj, parts[-1]


def generators_of(sim):
    """Generators in bus order then model order -- the order the hot loop visits them
    (solvers.py:84-86)."""
    out = []
    for bi, bus in enumerate(sim.busses):
        for m in bus.models:
            if hasattr(m, 'differential'):
                out.append((bi, m))
    return out


def channel_names(sim):
    names = []
    for bi, g in generators_of(sim):
        loc = list(STATE_NAMES)
        if g.exciter:
            loc += AVR_NAMES
        if g.governor:
            loc += GOV_NAMES
        if g.stabilizer:
            loc += PSS_NAMES
        loc += ALG_NAMES
        names += ['%s.%s' % (g.name, n) for n in loc]
    for bus in sim.busses:
        names += ['bus:%s.v_re' % bus.name, 'bus:%s.v_im' % bus.name]
    return names


def record_everything(sim):
    """Record function that captures every state, every algebraic the reference keeps on the
    machine object, and every bus voltage.  Called by run() after each step
    (simulator.py:540-543)."""
    import torch
    P = sim.parallel_sims

    def col(v):
        if not torch.is_tensor(v):
            v = torch.ones((P, 1), dtype=torch.float64) * v
        v = v.real if v.is_complex() else v
        return v.reshape(P, 1).to(torch.float64)

    out = []
    for bi, g in generators_of(sim):
        x = g.get_state_vector()            # [P, s, 1]  synchronous_machine.py:230-246
        for k in range(x.shape[1]):
            out.append(col(x[:, k]))
        out += [col(g.p_e), col(g.i_d), col(g.i_q), col(g.v_bb.real), col(g.v_bb.imag),
                col(g.e_fd), col(g.p_m)]
    for bus in sim.busses:
        out += [col(bus.voltage.real), col(bus.voltage.imag)]
    return out


def build(grid, P, solver, dt, sim_time, events=(), overrides=None, requires_grad=False,
          backend='torch', param_events=()):
    """Build a reference simulation.  ``overrides`` = {('gen', name, attr): array[P]} replaces
    model attributes by complex128 [P,1] tensors exactly like examples/models/ibb_model/ibb_opt.py:48-57."""
    simulator = _import_reference(backend)
    import numpy as np
    os.environ.pop('DIFFPSSI_FORCE_INTEGRATOR', None)
    sim = simulator.PowerSystemSimulation(time_step=dt, sim_time=sim_time, parallel_sims=P,
                                          solver=solver, grid_data=grid, verbose=False)
    for (t0, t1, bus) in events:
        sim.add_sc_event(t0, t1, bus)
    leaves = {}
    if overrides:
        if backend == 'torch':
            import torch
        for path, val in overrides.items():
            obj, attr = resolve(sim, path)
            old = getattr(obj, attr)
            arr = np.asarray(val, dtype=np.float64).reshape(P, 1)
            if backend == 'torch':
                is_c = torch.is_tensor(old) and old.is_complex()
                # the reference examples always hand in complex128 leaves, also for float64 block parameters
                # (k2a_opt.py:111-185); honour the attribute's own dtype unless a grad is wanted.
                # limiter bounds feed torch.minimum/maximum (blocks.py:181, 241), which reject complex tensors
                is_limit = attr in ('limit', 'lim_min', 'lim_max')
                want_grad = requires_grad and not is_limit
                dt_ = torch.complex128 if (is_c or want_grad) else torch.float64
                t = torch.tensor(arr, dtype=dt_, requires_grad=want_grad)
                if is_limit:
                    setattr(obj, attr, t)
                    continue
            else:
                t = arr.astype(np.complex128) if np.iscomplexobj(old) else arr
            setattr(obj, attr, t)
            leaves[path] = t
    # parameter events (simulator.py:464-474): (time, ('gen', name, attr), value) -> setattr at the first t >= time
    for (t_ev, path, value) in param_events:
        obj, attr = resolve(sim, path)
        old = getattr(obj, attr)
        if backend == 'torch':
            import torch
            dt_ = old.dtype if torch.is_tensor(old) else torch.float64
            new = torch.ones((P, 1), dtype=dt_) * value
        else:
            new = np.ones((P, 1)) * value
        sim.add_param_event(t_ev, obj, attr, new)
    return sim, leaves


def init_snapshot(sim):
    """Snapshot everything the hot loop consumes right after the reference's initialize()
    (simulator.py:438-450): x0 per generator, the held constants, Z = inv(Y_dyn), bus voltages."""
    import numpy as np

    def npy(v):
        v = v.detach() if hasattr(v, 'detach') else v
        return np.asarray(v)

    snap = {'Z0': npy(sim.inverse_dynamic_y_matrix)}
    for bi, g in generators_of(sim):
        pre = 'gen:%s.' % g.name
        snap[pre + 'x0'] = npy(g.get_state_vector()).real[:, :, 0]
        snap[pre + 'p_m'] = npy(g.p_m).real.reshape(-1)
        snap[pre + 'e_fd'] = npy(g.e_fd).real.reshape(-1)
        if g.exciter:
            snap[pre + 'avr_bias'] = npy(g.exciter.bias).real.reshape(-1)
            snap[pre + 'avr_vset'] = npy(g.exciter.v_setpoint).real.reshape(-1)
        if g.governor:
            snap[pre + 'gov_pref'] = npy(g.governor.p_ref).real.reshape(-1)
    snap['bus_v'] = np.stack([npy(b.voltage).reshape(-1) for b in sim.busses], axis=1)
    return snap


def run_full(sim, record=None):
    """run() (simulator.py:497-552) with the record-everything function (or ``record``).
    Returns (t, rec[P,T,C] float64 ndarray, names, snapshot-after-initialize, rec tensor)."""
    import numpy as np
    snap = {}
    orig_init = sim.initialize

    def hooked():
        orig_init()
        snap.update(init_snapshot(sim))

    sim.initialize = hooked
    sim.set_record_function(record or record_everything)
    try:
        t, rec_t = sim.run()
    finally:
        sim.initialize = orig_init
    rec = rec_t.detach().numpy() if hasattr(rec_t, 'detach') else np.asarray(rec_t)
    names = channel_names(sim) if record is None else None
    return np.asarray(t), np.ascontiguousarray(rec.real), names, snap, rec_t
