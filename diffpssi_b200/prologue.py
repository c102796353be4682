"""Host prologue: everything ``run()`` does *before* the time loop, in torch fp64/complex128.

This is the once-per-run part of the reference (SURVEY.md section 8b draws the kernel boundary after it):

    admittance assembly      simulator.py:350-426   (+ branch formulas static_models.py:296-312, 397-413)
    Newton-Raphson load flow load_flow.py:9-179
    model initialisation     synchronous_machine.py:279-323, exciters.py:129-138, governors.py:136-145,
                             stabilizers.py:135-145, blocks.py:196-207/332-342/426-436, static_models.py:515-527
    Z = inv(Y_dyn), V = Z I  simulator.py:386, 447-450

It is written with differentiable torch ops so the autograd graph from user leaf tensors
(``gen.h = torch.tensor(..., requires_grad=True)``) to the kernel inputs exists; the step kernel's
backward hands gradients w.r.t. those inputs back to this graph.

Batch handling.  Every quantity is a ``[B,1]`` tensor with ``B in {1, P}``: attributes that are identical in
every sim are collapsed to one row (broadcasting does the rest), so the load flow and the matrix inverse run
once when the network is shared by the batch and per sim when somebody varies ``x_d_st``, a load, ...
The collapse doubles as the kernel's uniform / per-sim classification of its inputs.

Two reference quirks are inherited deliberately because parity is judged against the reference's torch path:
  * line / transformer parameters are read from the **last** sim only.  The reference turns bus indices into
    ``[P,1]`` index tensors (static_models.py:321-322), so ``Y[:, from, to] += v`` (simulator.py:367-380)
    scatters with duplicate indices and the last row wins for every sim (and costs O(P^2) there).
  * the load flow stops when ``max |mismatch| < 1e-8`` over the *whole batch*, after applying one more
    update (load_flow.py:147-167).
"""
import math

import torch

from .components import SynchMachine, Load, Shunt, known_uniform

C128 = torch.complex128
F64 = torch.float64

# ---- kernel-level per-generator constants ("slots"); order is part of the C ABI (include/dps.h) -------------
SLOTS = (
    # machine (synchronous_machine.py:158-164, 256-277)
    'inv_2h', 'd', 'inv_td0t', 'xd_xdt', 'inv_tq0t', 'xq_xqt', 'inv_td0st', 'xdt_xdst', 'inv_tq0st', 'xqt_xqst',
    'inv_xdst', 'inv_xqst', 'snr', 'pm0', 'efd0',
    # controller constants derived from the machine's initial state (exciters.py:137-138, simulator.py:319,
    # governors.py:145): per-sim whenever a machine parameter is
    'a_vset', 'a_bias', 'g_pref',
    # SEXS (exciters.py:102-115; blocks.py:143-150, 290, 319)
    'a_r', 'a_omr', 'a_inv_tb', 'a_inv_te', 'a_k', 'a_emin', 'a_emax',
    # TGOV1 (governors.py:105-123)
    'g_inv_t1', 'g_gain', 'g_vmin', 'g_vmax', 'g_r', 'g_omr', 'g_inv_t3', 'g_dt',
    # STAB1 (stabilizers.py:106-121; blocks.py:382, 413, 241)
    's_inv_tw', 's_kw', 's_r1', 's_omr1', 's_inv_t3', 's_r2', 's_omr2', 's_inv_t4', 's_hlim',
)
N_REG_SLOTS = 18        # leading slots the fast kernel keeps in registers (DPS_N_REG_SLOTS)
SLOT_INDEX = {n: i for i, n in enumerate(SLOTS)}
N_SLOTS = len(SLOTS)

CTRL_AVR, CTRL_GOV, CTRL_PSS = 1, 2, 4


class _Cols(object):
    """Fetches model attributes as ``[B,1]`` tensors on the run device, collapsing batch-uniform ones."""

    def __init__(self, P, device, cache=None):
        self.P, self.device = P, device
        self._dev = torch.device(device)
        # id(tensor) -> (tensor, version): tensors already found batch-uniform.  Keeps run() from re-checking (one
        # device sync each) the hundreds of default parameter tensors that never change between runs.
        self.cache = cache if cache is not None else {}

    def __call__(self, value, dtype=None, collapse=True, last_row=False):
        P = self.P
        if not torch.is_tensor(value):
            dt = dtype or (C128 if isinstance(value, complex) else F64)
            return torch.full((1, 1), value, dtype=dt, device=self.device)
        t = value
        if t.device != self._dev and t.device.type != self._dev.type:
            t = t.to(self.device)
        elif t.device != self._dev and self._dev.index is not None and t.device.index != self._dev.index:
            t = t.to(self.device)
        if t.dim() == 0:
            t = t.reshape(1, 1)
        elif t.dim() == 1:
            t = t.reshape(-1, 1)
        if last_row:
            t = t[-1:]
        elif collapse and t.shape[0] == P and P > 1 and not t.requires_grad:
            # (collapse BEFORE any dtype conversion: converting one row instead of P)
            hit = self.cache.get(id(value))
            if (hit is not None and hit[0] is value and hit[1] == value._version) or known_uniform(value):
                t = t[:1]
            elif bool((t == t[:1]).all()):
                self.cache[id(value)] = (value, value._version)
                t = t[:1]
        if dtype is not None and t.dtype != dtype:
            if t.is_complex() and not dtype.is_complex:
                t = t.real
            t = t.to(dtype)
        return t


def _real(t):
    return t.real if t.is_complex() else t


def generators_of(sim):
    """Generators in the order the reference's solver visits them: bus order, then model order
    (solvers.py:84-86)."""
    out = []
    for bi, bus in enumerate(sim.busses):
        for m in bus.models:
            if isinstance(m, SynchMachine):
                out.append((bi, m))
    return out


# ==================================================================================================
# admittance matrices
# ==================================================================================================
def _branch_terms(sim, col):
    """(from, to, y_off, y_diag) per branch, in the reference's accumulation order: lines, then transformers."""
    terms = []
    for ln in sim.lines:
        r, x, b = (col(v, F64, last_row=True) for v in (ln.r, ln.x, ln.b))
        series = 1 / (r + 1j * x)
        terms.append((ln.from_bus_id, ln.to_bus_id, -series, series + 1j * b / 2))
    for tr in sim.trafos:
        r, x = (col(v, F64, last_row=True) for v in (tr.r, tr.x))
        series = 1 / (r + 1j * x)
        ratio_num, ratio_den = tr.s_n, tr.s_n_sys
        terms.append((tr.from_bus_id, tr.to_bus_id,
                      -series * ratio_num / ratio_den, (series + 1j * tr.b / 2) * ratio_num / ratio_den))
    return terms


def _assemble(sim, col, diag_contrib):
    """Y[B,n,n] = branches + per-bus diagonal contributions (simulator.py:362-384 / 400-422).
    Built out-of-place so autograd sees every contribution."""
    n = len(sim.busses)
    entries = {}

    def add(i, j, v):
        v = v.reshape(-1).to(C128)
        entries[(i, j)] = v if (i, j) not in entries else entries[(i, j)] + v

    for f, t, y_off, y_diag in _branch_terms(sim, col):
        add(f, t, y_off)
        add(t, f, y_off)
        add(f, f, y_diag)
        add(t, t, y_diag)
    for i, contribs in enumerate(diag_contrib):
        for v in contribs:
            add(i, i, v)
    B = max([v.shape[0] for v in entries.values()] + [1])
    keys = sorted(entries)
    flat = torch.tensor([i * n + j for i, j in keys], dtype=torch.long, device=col.device)
    vals = torch.stack([entries[k].expand(B) for k in keys], dim=1)                       # [B, nnz]
    y = torch.zeros((B, n * n), dtype=C128, device=col.device).index_copy(1, flat, vals)
    return y.reshape(B, n, n)


def static_admittance(sim, col):
    """Load-flow admittance: branches + shunts (generators and loads contribute zero, static_models.py:497-513,
    synchronous_machine.py:337-338)."""
    diag = []
    for bus in sim.busses:
        diag.append([col(m.y_shunt, C128) for m in bus.models if isinstance(m, Shunt)])
    return _assemble(sim, col, diag)


def dynamic_admittance(sim, col):
    """Dynamic admittance: branches + shunts + constant-Z loads + generator Norton admittances
    ``-j/x_d'' * S_n/S_sys`` (synchronous_machine.py:336)."""
    diag = []
    for bus in sim.busses:
        c = []
        for m in bus.models:
            if isinstance(m, Shunt):
                c.append(col(m.y_shunt, C128))
            elif isinstance(m, Load):
                c.append(m.y_load)
            elif isinstance(m, SynchMachine):
                c.append(-1j / col(m.x_d_st, C128) * (col(m.s_n, C128) / m.s_n_sys))
        diag.append(c)
    return _assemble(sim, col, diag)


# ==================================================================================================
# load flow
# ==================================================================================================
def load_flow(sim, col, y_bus, max_iter=600, tol=1e-8):
    """Polar Newton-Raphson power flow, batched over B.  Same unknown ordering, update sequence and stopping
    rule as load_flow.py:37-170.  Returns (V[B,n], S_calc[B,n], iterations)."""
    n = len(sim.busses)
    dev = col.device
    pv = [i for i, b in enumerate(sim.busses) if b.lf_type == 'PV']
    pq = [i for i, b in enumerate(sim.busses) if b.lf_type == 'PQ']
    ang_ids = pv + pq                                  # angle unknowns / P equations
    n_ang = len(ang_ids)

    # scheduled injections (static_models.py:165-176; synchronous_machine.py:347; static_models.py:495)
    s_parts = []
    for bus in sim.busses:
        s = torch.zeros((1, 1), dtype=C128, device=dev)
        for m in bus.models:
            if isinstance(m, SynchMachine):
                s = s + (col(m.p_soll_mw, C128) + 1j * 0) / m.s_n_sys
            elif isinstance(m, Load):
                s = s + (-(col(m.p_soll_mw, F64) + 1j * col(m.q_soll_mvar, F64)) / m.s_n_sys)
        s_parts.append(s)
    v_parts = [col(b.voltage, C128) for b in sim.busses]
    B = max([y_bus.shape[0]] + [s.shape[0] for s in s_parts] + [v.shape[0] for v in v_parts])
    s_soll = torch.cat([s.expand(B, 1) for s in s_parts], dim=1)              # [B,n]
    V = torch.cat([v.expand(B, 1) for v in v_parts], dim=1)                   # [B,n]
    G, Bm = y_bus.real, y_bus.imag
    ia = torch.tensor(ang_ids, device=dev, dtype=torch.long)
    iq = torch.tensor(pq, device=dev, dtype=torch.long)
    it = 0
    while True:
        s_calc = V * torch.conj(torch.matmul(y_bus, V.unsqueeze(-1)).squeeze(-1))
        th, vm = torch.angle(V), torch.abs(V)
        dth = th.unsqueeze(2) - th.unsqueeze(1)                               # theta_i - theta_k
        sin, cos = torch.sin(dth), torch.cos(dth)
        a = G * sin - Bm * cos
        c = G * cos + Bm * sin
        vi, vk = vm.unsqueeze(2), vm.unsqueeze(1)
        Pc, Qc = s_calc.real, s_calc.imag
        Gd = torch.diagonal(G, dim1=1, dim2=2)
        Bd = torch.diagonal(Bm, dim1=1, dim2=2)
        # J holds minus the usual Jacobian, exactly as load_flow.py:75-140 builds it
        J11 = -vi * vk * a
        J11 = torch.where(torch.eye(n, dtype=torch.bool, device=dev), (vm ** 2 * Bd + Qc).unsqueeze(2).expand(B, n, n), J11)
        J12 = -vi * c
        J12 = torch.where(torch.eye(n, dtype=torch.bool, device=dev), (-vm * Gd - Pc / vm).unsqueeze(2).expand(B, n, n), J12)
        J21 = vi * vk * c
        J21 = torch.where(torch.eye(n, dtype=torch.bool, device=dev), (vm ** 2 * Gd - Pc).unsqueeze(2).expand(B, n, n), J21)
        J22 = -vi * a
        J22 = torch.where(torch.eye(n, dtype=torch.bool, device=dev), (vm * Bd - Qc / vm).unsqueeze(2).expand(B, n, n), J22)
        top = torch.cat([J11[:, ia][:, :, ia], J12[:, ia][:, :, iq]], dim=2)
        bot = torch.cat([J21[:, iq][:, :, ia], J22[:, iq][:, :, iq]], dim=2)
        J = torch.cat([top, bot], dim=1)
        err = s_calc - s_soll
        rhs = torch.cat([err.real[:, ia], err.imag[:, iq]], dim=1).unsqueeze(-1)
        dx = torch.linalg.solve(-J, rhs).squeeze(-1)                           # [B, n_ang + n_pq]
        # update: angles of PV and PQ buses, then magnitudes of PQ buses (load_flow.py:150-158)
        va = V[:, ia]
        V = V.index_copy(1, ia, torch.abs(va) * torch.exp(1j * (torch.angle(va) - dx[:, :n_ang])))
        vq = V[:, iq]
        V = V.index_copy(1, iq, (torch.abs(vq) - dx[:, n_ang:]) * torch.exp(1j * torch.angle(vq)))
        worst = float(torch.max(torch.abs(rhs)))
        if worst < tol:
            break
        if it > max_iter:
            raise ValueError('Load flow did not converge')
        it += 1
    s_calc = V * torch.conj(torch.matmul(y_bus, V.unsqueeze(-1)).squeeze(-1))
    return V, s_calc, it


# ==================================================================================================
# model initialisation -> kernel inputs
# ==================================================================================================
class InitResult(object):
    """What the time loop consumes."""
    __slots__ = ('gens', 'gen_bus', 'ctrl', 'x0', 'slots', 'xst', 'Z0', 'Ydiag', 'bus_v', 's_calc', 'lf_iters', 'inj0')


def _clamp(x, lo, hi):
    return torch.minimum(torch.maximum(x, lo), hi)


def _fingerprint(values):
    """(key, keep-alive list, any requires_grad) of a list of model attributes: tensors are identified by object identity
    and version counter (in-place writes bump it, rebinding changes the identity), plain numbers by value."""
    key, keep, grad = [], [], False
    for v in values:
        if torch.is_tensor(v):
            key.append((id(v), v._version))
            keep.append(v)                       # a live reference: the id cannot be recycled while the entry exists
            grad = grad or v.requires_grad
        else:
            key.append(v)
    return tuple(key), keep, grad


def _network_inputs(sim):
    """Every attribute the static admittance matrix and the load flow read."""
    vals = []
    for ln in sim.lines:
        vals += [ln.from_bus_id, ln.to_bus_id, ln.r, ln.x, ln.b]
    for tr in sim.trafos:
        vals += [tr.from_bus_id, tr.to_bus_id, tr.r, tr.x, tr.b, tr.s_n, tr.s_n_sys]
    for bus in sim.busses:
        vals += [bus.lf_type, bus.voltage]
        for m in bus.models:
            if isinstance(m, Shunt):
                vals += ['shunt', m.y_shunt]
            elif isinstance(m, SynchMachine):
                vals += ['gen', m.p_soll_mw, m.s_n_sys]
            elif isinstance(m, Load):
                vals += ['load', m.p_soll_mw, m.q_soll_mvar, m.s_n_sys]
    return vals


def _dynamic_inputs(sim):
    """What the dynamic admittance matrix reads on top of :func:`_network_inputs` (the generator Norton admittances)."""
    vals = []
    for _, g in generators_of(sim):
        vals += [g.x_d_st, g.s_n, g.s_n_sys]
    return vals


def initialize(sim, device):
    """Load flow + model initialisation + Z = inv(Y_dyn) + initial bus voltages.
    Mirrors simulator.py:438-450.  Everything returned is ``[B,1]`` (B in {1,P}) real float64, except Z0/Ydiag/bus_v.

    The load flow and ``inv(Y_dyn)`` are pure functions of model attributes that an identification loop does not touch
    between runs (it varies H, D, X_d', ... per trajectory): their results are memoised on the simulation object, keyed
    by identity + version of every input tensor, and recomputed whenever any input changed or requires grad."""
    P = sim.parallel_sims
    col = _Cols(P, device, getattr(sim, '_uniform_cache', None))
    res = InitResult()
    cache = sim.__dict__.setdefault('_prologue_cache', {})
    loads = [m for bus in sim.busses for m in bus.models if isinstance(m, Load)]
    lf_key, lf_keep, lf_grad = _fingerprint(_network_inputs(sim))
    hit = cache.get('lf')
    if hit is not None and not lf_grad and hit['key'] == lf_key:
        V, s_calc, iters = hit['out']
        lf_cached = True
    else:
        y_static = static_admittance(sim, col)
        V, s_calc, iters = load_flow(sim, col, y_static)
        lf_cached = False
    res.lf_iters = iters
    res.s_calc = s_calc

    gens = generators_of(sim)
    res.gens = [g for _, g in gens]
    res.gen_bus = [bi for bi, _ in gens]
    res.ctrl, res.x0, res.slots, res.xst = [], [], [], []

    # loads become constant admittances from the load-flow voltage (static_models.py:515-527)
    if lf_cached:
        for m, y in zip(loads, hit['y_load']):
            m.y_load = y
    else:
        for bi, bus in enumerate(sim.busses):
            v_bb = V[:, bi:bi + 1]
            for m in bus.models:
                if isinstance(m, Load):
                    s_load = (col(m.p_soll_mw, F64) + 1j * col(m.q_soll_mvar, F64)) / m.s_n_sys
                    z_load = torch.conj(torch.abs(v_bb) ** 2 / s_load)
                    m.y_load = 1 / z_load
        cache.pop('lf', None)
        cache.pop('dyn', None)
        if not lf_grad:
            cache['lf'] = {'key': lf_key, 'keep': lf_keep, 'out': (V, s_calc, iters), 'y_load': [m.y_load for m in loads]}

    # Fine-grained memoisation of the model initialisation: every derived quantity is a pure function of a few model
    # attributes (and of earlier derived quantities); it is recomputed only when one of those *objects* changed
    # (identity + version) or requires grad.  An identification loop replaces a handful of per-sim tensors per run
    # (H, D, X_d' ...), so ~90 % of the ~1500 small tensor ops of this loop are skipped from the second run on.
    mstore = cache.setdefault('memo', {})
    if not lf_cached:
        mstore.clear()

    def memo(tag, fn, *inputs):
        key, grad = [], False
        for t in inputs:
            if torch.is_tensor(t):
                key.append((id(t), t._version))
                grad = grad or t.requires_grad
            else:
                key.append(t)
        if grad:
            return fn(*inputs)
        key = tuple(key)
        hit = mstore.get(tag)
        if hit is not None and hit[0] == key:
            return hit[2]
        out = fn(*inputs)
        mstore[tag] = (key, inputs, out)        # `inputs` keeps the keyed objects alive (ids cannot be recycled)
        return out

    def cc(v):
        return col(v, C128)

    def cn(v):
        return col(v, None)

    for gi, (bi, g) in enumerate(gens):
        snr = memo((gi, 'snr'), lambda s_n, sys: cc(s_n) / sys, g.s_n, g.s_n_sys)

        def base(V_, S_, snr_, x_q, bi=bi):
            # synchronous_machine.py:290-314
            v_bb, s_bus = V_[:, bi:bi + 1], S_[:, bi:bi + 1]
            i_gen = torch.conj(s_bus / v_bb) / snr_
            e = v_bb + 1j * cc(x_q) * i_gen
            delta = torch.angle(e)
            rot = torch.exp(1j * (math.pi / 2 - delta))
            i_dq = i_gen * rot
            v_dq = v_bb * rot
            return (e, delta, i_dq.real, i_dq.imag, v_dq.real, v_dq.imag, s_bus.real / snr_, torch.zeros_like(delta),
                    torch.zeros_like(e), delta.to(C128) if not delta.is_complex() else delta)
        e, delta, i_d, i_q, v_d, v_q, p_m, zero_r, zero_c, delta_c = memo((gi, 'base'), base, V, s_calc, snr, g.x_q)
        e_q_t = memo((gi, 'e_q_t'), lambda vq, x, i: vq + cc(x) * i, v_q, g.x_d_t, i_d)
        e_d_t = memo((gi, 'e_d_t'), lambda vd, x, i: vd - cc(x) * i, v_d, g.x_q_t, i_q)
        e_q_st = memo((gi, 'e_q_st'), lambda vq, x, i: vq + cc(x) * i, v_q, g.x_d_st, i_d)
        e_d_st = memo((gi, 'e_d_st'), lambda vd, x, i: vd - cc(x) * i, v_d, g.x_q_st, i_q)
        e_fd = memo((gi, 'e_fd'), lambda eqt, i, xd, xdt: eqt + i * (cc(xd) - cc(xdt)), e_q_t, i_d, g.x_d, g.x_d_t)

        x0 = [zero_r, delta, _real(e_q_t), _real(e_d_t), _real(e_q_st), _real(e_d_st)]
        inv = lambda tag, attr: memo((gi, tag), lambda t: _real(1 / cc(t)), attr)
        dif = lambda tag, a1, a2: memo((gi, tag), lambda u, v: _real(cc(u) - cc(v)), a1, a2)
        sl = {
            'inv_2h': memo((gi, 'inv_2h'), lambda h: _real(1 / (2 * cc(h))), g.h),
            'd': memo((gi, 'd'), lambda d: _real(cc(d)), g.d),
            'inv_td0t': inv('inv_td0t', g.t_d0_t), 'xd_xdt': dif('xd_xdt', g.x_d, g.x_d_t),
            'inv_tq0t': inv('inv_tq0t', g.t_q0_t), 'xq_xqt': dif('xq_xqt', g.x_q, g.x_q_t),
            'inv_td0st': inv('inv_td0st', g.t_d0_st), 'xdt_xdst': dif('xdt_xdst', g.x_d_t, g.x_d_st),
            'inv_tq0st': inv('inv_tq0st', g.t_q0_st), 'xqt_xqst': dif('xqt_xqst', g.x_q_t, g.x_q_st),
            'inv_xdst': inv('inv_xdst', g.x_d_st), 'inv_xqst': inv('inv_xqst', g.x_q_st),
            'snr': _real(snr), 'pm0': _real(p_m), 'efd0': _real(e_fd),
        }
        flags = 0
        if g.exciter:
            flags |= CTRL_AVR
            ex = g.exciter
            ll, pt = ex.lead_lag, ex.pt1_lim
            # blocks.py:206-207 then 341-342 via exciters.py:137-138
            x_pt = memo((gi, 'a_xpt'), lambda ef, lo, hi: _clamp(_real(ef), _real(cn(lo)), _real(cn(hi))),
                        e_fd, pt.lim_min, pt.lim_max)
            bias = memo((gi, 'a_bias'), lambda ef, k: ef / cn(k), e_fd, pt.gain_pt1)
            x0 += [_real(bias), x_pt]

            def avr_consts(t_a, t_b, t_e, k, lo, hi, vset):
                r = cn(t_a) / cn(t_b)
                return {'a_r': _real(r), 'a_omr': _real(1 - r), 'a_inv_tb': _real(1 / cn(t_b)),
                        'a_inv_te': _real(1 / cn(t_e)), 'a_k': _real(cn(k)), 'a_emin': _real(cn(lo)),
                        'a_emax': _real(cn(hi)), 'a_vset': _real(cn(vset))}
            sl.update(memo((gi, 'a_consts'), avr_consts, ll.t_1, ll.t_2, pt.t_pt1, pt.gain_pt1, pt.lim_min, pt.lim_max,
                           ex.v_setpoint))
            sl['a_bias'] = _real(bias)
            ex.bias = bias
        if g.governor:
            flags |= CTRL_GOV
            gv = g.governor
            pt, ll = gv.pt1_lim, gv.lead_lag
            # governors.py:144-145: lead-lag state = p_m; PT1 state = clamp(p_m); p_ref = p_m / gain
            x_pt = memo((gi, 'g_xpt'), lambda pm, lo, hi: _clamp(_real(pm), _real(cn(lo)), _real(cn(hi))),
                        p_m, pt.lim_min, pt.lim_max)
            p_ref = memo((gi, 'g_pref'), lambda pm, gain: pm / cn(gain), p_m, pt.gain_pt1)
            x0 += [x_pt, _real(p_m)]

            def gov_consts(t_1, gain, lo, hi, t_2, t_3, d_t):
                r = cn(t_2) / cn(t_3)
                return {'g_inv_t1': _real(1 / cn(t_1)), 'g_gain': _real(cn(gain)), 'g_vmin': _real(cn(lo)),
                        'g_vmax': _real(cn(hi)), 'g_r': _real(r), 'g_omr': _real(1 - r), 'g_inv_t3': _real(1 / cn(t_3)),
                        'g_dt': _real(cn(d_t))}
            sl.update(memo((gi, 'g_consts'), gov_consts, pt.t_pt1, pt.gain_pt1, pt.lim_min, pt.lim_max, ll.t_1, ll.t_2,
                           gv.d_t))
            sl['g_pref'] = _real(p_ref)
            gv.p_ref = p_ref
        if g.stabilizer:
            flags |= CTRL_PSS
            st = g.stabilizer
            # synchronous_machine.py:323 initialises the PSS with e_fd*0 -> all three states zero
            # (blocks.py:341, 435: washout state = k_w * 0)
            xw, z = memo((gi, 's_x0'), lambda ef, k_w: (_real(cn(k_w) * (ef * 0)), _real(ef) * 0), e_fd, st.washout.k_w)
            x0 += [xw, z, z]

            def pss_consts(k_w, t_w, t1, t3, t2, t4, lim):
                r1, r2 = cn(t1) / cn(t3), cn(t2) / cn(t4)
                return {'s_inv_tw': _real(1 / cn(t_w)), 's_kw': _real(cn(k_w)), 's_r1': _real(r1), 's_omr1': _real(1 - r1),
                        's_inv_t3': _real(1 / cn(t3)), 's_r2': _real(r2), 's_omr2': _real(1 - r2),
                        's_inv_t4': _real(1 / cn(t4)), 's_hlim': _real(cn(lim))}
            sl.update(memo((gi, 's_consts'), pss_consts, st.washout.k_w, st.washout.t_w, st.lead_lag1.t_1,
                           st.lead_lag1.t_2, st.lead_lag2.t_1, st.lead_lag2.t_2, st.limiter.limit))
        res.ctrl.append(flags)
        res.xst.append(memo((gi, 'xst'), lambda u, v: (cc(u), cc(v)), g.x_d_st, g.x_q_st))
        res.x0.append(x0)
        res.slots.append(sl)
        # expose initial values on the container, as the reference does
        g.omega, g.delta = zero_c, delta_c
        g.e_q_t, g.e_d_t, g.e_q_st, g.e_d_st = e_q_t, e_d_t, e_q_st, e_d_st
        g.p_m, g.e_fd = p_m, e_fd

    dyn_key, dyn_keep, dyn_grad = _fingerprint(_dynamic_inputs(sim))
    hit = cache.get('dyn')
    if hit is not None and lf_cached and not dyn_grad and hit['key'] == dyn_key:
        res.Ydiag, res.Z0 = hit['out']
    else:
        y_dyn = dynamic_admittance(sim, col)
        res.Ydiag = torch.diagonal(y_dyn, dim1=1, dim2=2)
        res.Z0 = torch.linalg.inv(y_dyn)
        cache.pop('dyn', None)
        if 'lf' in cache and not dyn_grad:
            cache['dyn'] = {'key': dyn_key, 'keep': dyn_keep, 'out': (res.Ydiag, res.Z0)}
    res.bus_v = V
    return res


def initial_injections(res):
    """Norton current injections of every generator at x0 (synchronous_machine.py:256-260), [B,G] complex."""
    cur = []
    for x0, sl, (xdst, xqst) in zip(res.x0, res.slots, res.xst):
        delta, eq, ed = x0[1], x0[4], x0[5]
        i_d = eq / (1j * xdst) * torch.exp(1j * delta)
        i_q = ed / (1j * xqst) * torch.exp(1j * (delta - math.pi / 2))
        cur.append((i_d + i_q) * sl['snr'])
    B = max(c.shape[0] for c in cur)
    return torch.cat([c.expand(B, 1) for c in cur], dim=1)


# ==================================================================================================
# parameter events
# ==================================================================================================
INIT_DERIVED = ('pm0', 'efd0', 'a_vset', 'a_bias', 'g_pref')


def refresh_loop_slots(sim, init, device):
    """Constant slots after a ParamEvent changed a model attribute mid-run (static_models.py:82-95, applied at
    simulator.py:533-534).  The reference only ``setattr``s: nothing is re-initialised, so the values initialize() left
    behind (p_m / e_fd of uncontrolled machines, v_setpoint, bias, p_ref) and Z stay as they are; every other constant
    is re-derived from the current attributes with the same expressions as in :func:`initialize`."""
    col = _Cols(sim.parallel_sims, device, getattr(sim, '_uniform_cache', None))
    out = []
    for g, old in zip(init.gens, init.slots):
        c = {k: col(getattr(g, k), C128) for k in ('s_n', 'h', 'd', 'x_d', 'x_q', 'x_d_t', 'x_q_t', 'x_d_st', 'x_q_st',
                                                   't_d0_t', 't_q0_t', 't_d0_st', 't_q0_st')}
        sl = {
            'inv_2h': _real(1 / (2 * c['h'])), 'd': _real(c['d']),
            'inv_td0t': _real(1 / c['t_d0_t']), 'xd_xdt': _real(c['x_d'] - c['x_d_t']),
            'inv_tq0t': _real(1 / c['t_q0_t']), 'xq_xqt': _real(c['x_q'] - c['x_q_t']),
            'inv_td0st': _real(1 / c['t_d0_st']), 'xdt_xdst': _real(c['x_d_t'] - c['x_d_st']),
            'inv_tq0st': _real(1 / c['t_q0_st']), 'xqt_xqst': _real(c['x_q_t'] - c['x_q_st']),
            'inv_xdst': _real(1 / c['x_d_st']), 'inv_xqst': _real(1 / c['x_q_st']),
            'snr': _real(c['s_n'] / g.s_n_sys),
        }
        if g.exciter:
            ll, pt = g.exciter.lead_lag, g.exciter.pt1_lim
            r = col(ll.t_1, None) / col(ll.t_2, None)
            sl.update({'a_r': _real(r), 'a_omr': _real(1 - r), 'a_inv_tb': _real(1 / col(ll.t_2, None)),
                       'a_inv_te': _real(1 / col(pt.t_pt1, None)), 'a_k': _real(col(pt.gain_pt1, None)),
                       'a_emin': _real(col(pt.lim_min, None)), 'a_emax': _real(col(pt.lim_max, None))})
        if g.governor:
            pt, ll = g.governor.pt1_lim, g.governor.lead_lag
            r = col(ll.t_1, None) / col(ll.t_2, None)
            sl.update({'g_inv_t1': _real(1 / col(pt.t_pt1, None)), 'g_gain': _real(col(pt.gain_pt1, None)),
                       'g_vmin': _real(col(pt.lim_min, None)), 'g_vmax': _real(col(pt.lim_max, None)),
                       'g_r': _real(r), 'g_omr': _real(1 - r), 'g_inv_t3': _real(1 / col(ll.t_2, None)),
                       'g_dt': _real(col(g.governor.d_t, None))})
        if g.stabilizer:
            st = g.stabilizer
            r1 = col(st.lead_lag1.t_1, None) / col(st.lead_lag1.t_2, None)
            r2 = col(st.lead_lag2.t_1, None) / col(st.lead_lag2.t_2, None)
            sl.update({'s_inv_tw': _real(1 / col(st.washout.t_w, None)), 's_kw': _real(col(st.washout.k_w, None)),
                       's_r1': _real(r1), 's_omr1': _real(1 - r1), 's_inv_t3': _real(1 / col(st.lead_lag1.t_2, None)),
                       's_r2': _real(r2), 's_omr2': _real(1 - r2), 's_inv_t4': _real(1 / col(st.lead_lag2.t_2, None)),
                       's_hlim': _real(col(st.limiter.limit, None))})
        for k in INIT_DERIVED:
            if k in old:
                sl[k] = old[k]
        # the reference keeps p_m / e_fd as plain attributes of the machine: a ParamEvent may overwrite them too
        if not g.governor:
            sl['pm0'] = _real(col(g.p_m, C128))
        if not g.exciter:
            sl['efd0'] = _real(col(g.e_fd, C128))
        if g.exciter:
            sl['a_vset'] = _real(col(g.exciter.v_setpoint, None))
            sl['a_bias'] = _real(col(g.exciter.bias, None))
        if g.governor:
            sl['g_pref'] = _real(col(g.governor.p_ref, None))
        out.append(sl)
    return out
