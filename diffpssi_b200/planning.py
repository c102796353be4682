"""Host-side planning: turn (grid, initial state, events, recorder channels) into the flat buffers and the
``dps_topology`` the C ABI takes (include/dps.h).  Pure host logic -- testable without a GPU.
"""
import numpy as np
import torch

from . import engine
from .prologue import SLOTS, SLOT_INDEX, N_SLOTS, N_REG_SLOTS, CTRL_AVR, CTRL_GOV, CTRL_PSS

F64 = torch.float64
FIELD_VRE, FIELD_VIM = 16, 17
FAULT_ADMITTANCE = 1e6          # simulator.py:527


def event_schedule(time_grid, sc_events):
    """Per-step network configuration, reproducing simulator.py:524-531 *including* its multi-event quirk:

        for ev in sc_events:                      # list order
            if ev.is_active(t):  Z <- fault(Z_current, ev.bus)      (Y=inv(Z); Y[b,b]=1e6; Z=inv(Y))
            else:                Z <- Z_original

    Faulting is idempotent and order-independent (it *sets* diagonal entries of Y), so the matrix in force during a
    step is determined by a *set* of faulted buses: an inactive event wipes the set (also faults applied by earlier
    events in the list, and faults carried over from the previous step), an active one adds its bus.
    Returns (step_cfg uint8[T], cfg_sets) with cfg_sets[0] == () always.
    """
    cfg_sets = [()]
    step_cfg = np.zeros(len(time_grid), dtype=np.uint8)
    cur = frozenset()
    for k, t in enumerate(time_grid):
        for ev in sc_events:
            cur = (cur | {int(ev.bus)}) if ev.is_active(t) else frozenset()
        key = tuple(sorted(cur))
        if key not in cfg_sets:
            cfg_sets.append(key)
        step_cfg[k] = cfg_sets.index(key)
    if len(cfg_sets) > engine.MAX_CFG:
        raise engine.EngineError('too many distinct fault configurations (%d > %d)' % (len(cfg_sets), engine.MAX_CFG))
    if max(len(s) for s in cfg_sets) > engine.MAX_FAULT:
        raise engine.EngineError('too many simultaneously faulted buses')
    return step_cfg, cfg_sets


def canonical_x0(init, P):
    """[G,13,P] initial states in the kernel's canonical layout (absent controllers -> zero rows)."""
    rows = []
    for x0, flags in zip(init.x0, init.ctrl):
        it = iter(x0)
        mach = [next(it) for _ in range(6)]
        avr = [next(it), next(it)] if flags & CTRL_AVR else [None, None]
        gov = [next(it), next(it)] if flags & CTRL_GOV else [None, None]
        pss = [next(it), next(it), next(it)] if flags & CTRL_PSS else [None, None, None]
        dev = mach[1].device
        zero = torch.zeros(P, dtype=F64, device=dev)
        cols = [zero if v is None else v.reshape(-1).to(F64).expand(P) for v in mach + avr + gov + pss]
        rows.append(torch.stack(cols, dim=0))
    return torch.stack(rows, dim=0).contiguous()


def pack_slots(init, P, device, memo=None):
    """upar [G,42] (uniform constants), vpar [n_vary,P] (per-sim constants), slot_vary list[G*42].

    ``memo`` (a dict living on the simulation): upar is rebuilt only when one of the uniform slot tensors is a different
    object than in the previous run (the prologue memoises them, so in an identification loop they are not)."""
    zero1 = torch.zeros(1, dtype=F64, device=device)
    uniform, vrows, slot_vary, key = [], [], [], []
    for sl in init.slots:
        for name in SLOTS:
            v = sl.get(name)
            if v is None or v.shape[0] != 1 and v.numel() != 1:
                uniform.append(None)
                key.append(None)
                if v is None:
                    slot_vary.append(-1)
                else:
                    slot_vary.append(len(vrows))
                    vrows.append(v.reshape(-1).to(F64))
            else:
                uniform.append(v)
                key.append((id(v), v._version))
                slot_vary.append(-1)
    G = len(init.slots)
    hit = memo.get('upar') if memo is not None else None
    no_grad = not any(v is not None and v.requires_grad for v in uniform)
    if hit is not None and no_grad and hit[0] == key:
        upar = hit[2]
    else:
        upar = torch.cat([zero1 if v is None else v.reshape(-1).to(F64) for v in uniform]).reshape(G, N_SLOTS).contiguous()
        if memo is not None and no_grad:
            memo['upar'] = (key, uniform, upar)
    vpar = torch.stack(vrows, dim=0).contiguous() if vrows else None
    return upar, vpar, slot_vary


def common_controller_constants(init, upar, slot_vary, memo=None):
    """Host list [N_SLOTS] of the controller constants when every generator has all three controllers and their
    (batch-uniform) constants are identical between generators -- the stepper then takes them as constant-bank operands
    (``dps_topology.common_ctrl``).  None otherwise.

    The values are baked into the plan, so a controller constant that requires grad (it changes from one optimisation step
    to the next) rules the specialisation out.  ``memo`` (a dict living on the simulation) keeps the answer for as long as
    the controller-constant tensors are the same objects at the same version: no device->host read per run(), which also
    keeps run() capturable into a CUDA graph."""
    if any(f != (CTRL_AVR | CTRL_GOV | CTRL_PSS) for f in init.ctrl):
        return None
    G = len(init.ctrl)
    if any(slot_vary[g * N_SLOTS + k] >= 0 for g in range(G) for k in range(N_REG_SLOTS, N_SLOTS)):
        return None
    tensors = [sl[name] for sl in init.slots for name in SLOTS[N_REG_SLOTS:]]
    if any(t.requires_grad for t in tensors):
        return None
    key = tuple((id(t), t._version) for t in tensors)
    hit = memo.get('common_ctrl') if memo is not None else None
    if hit is not None and hit[0] == key:
        return hit[2]
    host = upar.detach().cpu()
    tail = host[:, N_REG_SLOTS:]
    out = [float(v) for v in host[0]] if bool((tail == tail[:1]).all()) else None
    if memo is not None:
        memo['common_ctrl'] = (key, tensors, out)        # `tensors` keeps the keyed objects alive
    return out


def common_controller_constants_keyed(ctrl, upar, slot_vary, key, memo):
    """:func:`common_controller_constants` for the kernel prologue: same answer, memoised on the identity + version of the
    raw controller parameters (``key``) instead of on slot tensors that are new views every run."""
    if any(f != (CTRL_AVR | CTRL_GOV | CTRL_PSS) for f in ctrl) or any(k is None for k in key):
        return None
    G = len(ctrl)
    if any(slot_vary[g * N_SLOTS + k] >= 0 for g in range(G) for k in range(N_REG_SLOTS, N_SLOTS)):
        return None
    hit = memo.get('common_ctrl_keyed')
    if hit is not None and hit[0] == key:
        return hit[1]
    host = upar.detach().cpu()
    tail = host[:, N_REG_SLOTS:]
    out = [float(v) for v in host[0]] if bool((tail == tail[:1]).all()) else None
    memo['common_ctrl_keyed'] = (key, out)
    return out


def observed_rows(init, channel_keys):
    """Rows of the reduced impedance matrix: the generator buses (generator order) then every recorded bus that hosts
    no generator.  Also rewrites recorder keys into (unique) C-ABI channels; ``rows[i]`` is the kernel recorder row
    that holds key i (two keys can name the same quantity, e.g. ``bus.voltage.real`` and ``gen.v_bb.real``)."""
    gen_bus = list(init.gen_bus)
    first_gen_on_bus = {}
    for gi, b in enumerate(gen_bus):
        first_gen_on_bus.setdefault(b, gi)
    obs_bus = list(gen_bus)
    chans, rows = [], []

    def channel(c):
        if c not in chans:
            chans.append(c)
        rows.append(chans.index(c))

    for key in channel_keys:
        if key[0] == 'gen':
            channel((engine.CH_GEN, key[1], key[2]))
        else:
            _, bus, part = key
            if bus in first_gen_on_bus:
                # a generator bus: its voltage is the v_bb the machine thread already holds
                channel((engine.CH_GEN, first_gen_on_bus[bus], FIELD_VIM if part else FIELD_VRE))
            else:
                if bus not in obs_bus[len(gen_bus):]:
                    obs_bus.append(bus)
                row = len(gen_bus) + obs_bus[len(gen_bus):].index(bus)
                channel((engine.CH_BUS_VIM if part else engine.CH_BUS_VRE, row, 0))
    if len(obs_bus) - len(gen_bus) > engine.MAX_EXTRA_OBS:
        raise engine.EngineError('too many recorded non-generator buses')
    return obs_bus, chans, rows


def split_complex_batch_last(z):
    """complex [B, ...] -> (re, im) float64 [..., B] contiguous (batch innermost, the kernels' layout)."""
    perm = list(range(1, z.dim())) + [0]
    return z.real.permute(perm).contiguous(), z.imag.permute(perm).contiguous()


def network_reduce_torch(Z0, Ydiag, obs_bus, gen_bus, cfg_sets, m_max, y_fault=FAULT_ADMITTANCE):
    """Differentiable torch twin of ``dps_network_reduce`` (csrc/dps_reduce.cuh), used instead of the kernel when a
    gradient has to flow into the network (``x_d_st``, ``s_n``, loads ... enter ``Y_dyn``).

    ``Z0`` complex [B,n,n], ``Ydiag`` complex [B,n] -> zsub float64 [n_cfg, n_obs+m_max, n_gen+m_max, 2, B] with the
    rows/columns beyond [n_obs, n_gen] (the kernel's scratch) left at zero.  Each faulted bus is the rank-1 update
    ``Z' = Z - f Z[:,b] Z[b,:]``, ``f = c/(1 + c Z_bb)``, ``c = y_fault - Y_bb`` of simulator.py:526-528."""
    B = max(Z0.shape[0], Ydiag.shape[0])
    rows, cols = len(obs_bus) + m_max, len(gen_bus) + m_max
    oi = torch.tensor(obs_bus, dtype=torch.long, device=Z0.device)
    gi = torch.tensor(gen_bus, dtype=torch.long, device=Z0.device)
    out = []
    for faulted in cfg_sets:
        Z = Z0
        for fb in faulted:
            c = y_fault - Ydiag[:, fb]
            f = c / (1 + c * Z[:, fb, fb])
            Z = Z - f[:, None, None] * Z[:, :, fb:fb + 1] * Z[:, fb:fb + 1, :]
        sub = Z[:, oi][:, :, gi].expand(B, len(obs_bus), len(gen_bus))            # [B, n_obs, n_gen]
        sub = torch.nn.functional.pad(torch.view_as_real(sub.resolve_conj()), (0, 0, 0, m_max, 0, m_max))
        out.append(sub.permute(1, 2, 3, 0))                                        # [rows, cols, 2, B]
    return torch.stack(out, dim=0).contiguous()


def build_topology(sim, init, channel_keys, solver_name, ckpt_stride=0):
    """-> (topo dict for engine.Plan, dict of host-side extras)."""
    ev_key = (len(sim.time), float(sim.time_step), tuple((e.start_time, e.end_time, int(e.bus)) for e in sim.sc_events))
    hit = sim.__dict__.get('_schedule_cache')
    if hit is None or hit[0] != ev_key:                # the per-step Python loop costs milliseconds: once per event list
        hit = (ev_key,) + event_schedule(sim.time, sim.sc_events)
        sim.__dict__['_schedule_cache'] = hit
    step_cfg, cfg_sets = hit[1], hit[2]
    obs_bus, chans, rows = observed_rows(init, channel_keys)
    G = len(init.gens)
    if G > engine.MAX_GW:
        raise engine.EngineError('%d generators: the stepper handles up to %d' % (G, engine.MAX_GW))
    m_max = max(len(s) for s in cfg_sets)
    topo = {
        'n_gen': G, 'solver': engine.SOLVER_ID[solver_name], 'n_steps': len(sim.time), 'dt': float(sim.time_step),
        'fn': float(sim.fn), 'ctrl': list(init.ctrl), 'n_obs': len(obs_bus),
        'z_rows': len(obs_bus) + m_max, 'z_cols': G + m_max, 'n_cfg': len(cfg_sets),
        'step_cfg': step_cfg.tolist(), 'z_per_sim': int(init.Z0.shape[0] > 1), 'rec': chans,
        'ckpt_stride': int(ckpt_stride),
    }
    extras = {'cfg_sets': cfg_sets, 'obs_bus': obs_bus, 'gen_bus': list(init.gen_bus), 'm_max': m_max,
              'key_rows': rows}
    return topo, extras
