"""TEST INFRASTRUCTURE -- feeds the C oracle from a ``diffpssi_b200`` simulation object.

Reads raw, reference-level parameters off the package's grid components and the initial state / full impedance
matrix off its prologue result, so that the oracle and the CUDA engine integrate *identical* inputs.  Used by tests/,
``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs only.
"""
import numpy as np
import torch

from . import oracle


def resolve_attr(sim, gen_name, attr):
    obj = sim.get_generator_by_name(gen_name)
    parts = attr.split('.')
    for p in parts[:-1]:
        obj = getattr(obj, p)
    return obj, parts[-1]


def _canonical_x0(init, P):
    """[G,13,P] canonical initial states (machine | avr | gov | pss, zeros for absent controllers)."""
    G = len(init.gens)
    out = np.zeros((G, oracle.N_STATES, P))
    for gi, (x0, flags) in enumerate(zip(init.x0, init.ctrl)):
        vals = [np.broadcast_to(v.detach().cpu().numpy().reshape(-1), (P,)) for v in x0]
        out[gi, :6] = vals[:6]
        off = 6
        if flags & oracle.CTRL_AVR:
            out[gi, 6:8] = vals[off:off + 2]
            off += 2
        if flags & oracle.CTRL_GOV:
            out[gi, 8:10] = vals[off:off + 2]
            off += 2
        if flags & oracle.CTRL_PSS:
            out[gi, 10:13] = vals[off:off + 3]
    return out


def oracle_inputs_from_sim(sim, init):
    """Raw reference-level parameters, x0 and full Z0 for oracle.run(), read off the package's own objects."""
    P = sim.parallel_sims
    G = len(init.gens)
    raw = np.zeros((G, oracle.N_RAW, P))

    def val(v):
        if torch.is_tensor(v):
            v = v.detach().cpu()
            v = (v.real if v.is_complex() else v).numpy().reshape(-1)
            return v if v.shape[0] == P else np.full(P, v[0])
        return np.full(P, float(v))

    for gi, g in enumerate(init.gens):
        def put(name, v):
            raw[gi, oracle.RAW_INDEX[name]] = val(v)
        for attr in ('h', 'd', 'x_d', 'x_q', 'x_d_t', 'x_q_t', 'x_d_st', 'x_q_st', 't_d0_t', 't_q0_t', 't_d0_st',
                     't_q0_st', 's_n'):
            put(attr, getattr(g, attr))
        put('s_n_sys', g.s_n_sys)
        put('p_m0', init.slots[gi]['pm0'])
        put('e_fd0', init.slots[gi]['efd0'])
        for attr, rn in oracle.ATTR_TO_RAW.items():
            if '.' not in attr:
                continue
            head = attr.split('.')[0]
            if getattr(g, head) is None:
                continue
            obj, name = resolve_attr(sim, g.name, attr)
            put(rn, getattr(obj, name))
        if g.exciter:
            put('a_v_set', init.slots[gi]['a_vset'])
            put('a_bias', init.slots[gi]['a_bias'])
        if g.governor:
            put('g_p_ref', init.slots[gi]['g_pref'])
    x0 = _canonical_x0(init, P)
    z0 = init.Z0.detach().cpu().numpy()
    events = [(e.start_time, e.end_time, e.bus) for e in sim.sc_events]
    return dict(n_bus=len(sim.busses), gen_bus=list(init.gen_bus), ctrl=list(init.ctrl), raw=raw, x0=x0, z0=z0,
                events=events, fn=sim.fn)


def run_oracle_for(sim, init, rec_channels=None, full=True, threads=None):
    inp = oracle_inputs_from_sim(sim, init)
    return oracle.run(inp['n_bus'], inp['gen_bus'], inp['ctrl'], sim.solver, float(sim.time_step), inp['fn'],
                      sim.time, inp['events'], inp['raw'], inp['x0'], inp['z0'], rec=rec_channels or (), full=full,
                      busv=full, threads=threads)
