"""initialize() on the device: the kernel path of the once-per-run prologue.

``prologue.initialize`` evaluates load flow, model initialisation and ``Z = inv(Y_dyn)`` as ~2000 small differentiable
torch operations (bit-identical to the reference, needed whenever a gradient has to flow through them).  A run that needs
no such gradient goes through three kernels of ``libdps_b200.so`` instead (``csrc/dps_prologue.cu``):

    dps_load_flow        load_flow.py:9-179           one block per batch row: mismatch, Jacobian, dense solve, update
    dps_dynamic_inverse  static_models.py:515-527,    loads -> admittances at the load-flow voltage, Z0 = inv(Y_dyn)
                         simulator.py:386, 400-422
    dps_model_init       synchronous_machine.py:279-323 + controller initialisations: x0, constant slots, Norton currents

The admittance matrices themselves (branches, shunts, generator Norton admittances) are assembled by a handful of small
matrix products with 0/1 incidence matrices (``_Network``) -- ``[B,n,n]`` with B = 1 when the network is shared -- and
memoised like the torch path's.  Everything returned has the structure of
``prologue.InitResult`` (``[B,1]`` views with B in {1, P}; which quantities are per-trajectory follows from which raw
parameters are), so planning, packing, recorder tracing and the oracle harness do not care which path produced it.
Agreement with the torch prologue: <= 1e-12 on every state, constant and matrix entry (tests/test_cuda_prologue.py); the
dense solves pivot like LAPACK but sum in a different order.
"""
import ctypes

import torch

from . import engine, prologue
from .components import SynchMachine, Load, Shunt
from .prologue import C128, F64, CTRL_AVR, CTRL_GOV, CTRL_PSS, SLOTS, SLOT_INDEX

MAX_BUS = 64                      # PRO_MAX_BUS of csrc/dps_prologue.cu

RAW = ('s_n', 's_n_sys', 'x_d', 'x_q', 'x_d_t', 'x_q_t', 'x_d_st', 'x_q_st', 't_d0_t', 't_q0_t', 't_d0_st', 't_q0_st', 'h', 'd',
       'a_t1', 'a_t2', 'a_te', 'a_k', 'a_emin', 'a_emax', 'a_vset',
       'g_t1', 'g_gain', 'g_vmin', 'g_vmax', 'g_t2', 'g_t3', 'g_dt',
       's_kw', 's_tw', 's_ll1_t1', 's_ll1_t2', 's_ll2_t1', 's_ll2_t2', 's_hlim')
N_RAW = len(RAW)
R = {n: i for i, n in enumerate(RAW)}

# which raw parameters (and 'base' = load-flow voltage / power of the generator bus, S_n, X_q) every output depends on:
# an output differs per trajectory iff one of its inputs does (the torch path gets the same answer from broadcasting)
_PM = ('base', 's_n', 's_n_sys')               # p_m = Re(S) / (S_n / S_sys)
_BASE = _PM + ('x_q',)                         # delta, i_d, i_q, v_d, v_q: e = v + j x_q i
_EFD = _BASE + ('x_d_t', 'x_d')
SLOT_DEPS = {
    'inv_2h': ('h',), 'd': ('d',), 'inv_td0t': ('t_d0_t',), 'xd_xdt': ('x_d', 'x_d_t'), 'inv_tq0t': ('t_q0_t',),
    'xq_xqt': ('x_q', 'x_q_t'), 'inv_td0st': ('t_d0_st',), 'xdt_xdst': ('x_d_t', 'x_d_st'), 'inv_tq0st': ('t_q0_st',),
    'xqt_xqst': ('x_q_t', 'x_q_st'), 'inv_xdst': ('x_d_st',), 'inv_xqst': ('x_q_st',), 'snr': ('s_n', 's_n_sys'),
    'pm0': _PM, 'efd0': _EFD, 'a_vset': ('a_vset',), 'a_bias': _EFD + ('a_k',), 'g_pref': _PM + ('g_gain',),
    'a_r': ('a_t1', 'a_t2'), 'a_omr': ('a_t1', 'a_t2'), 'a_inv_tb': ('a_t2',), 'a_inv_te': ('a_te',), 'a_k': ('a_k',),
    'a_emin': ('a_emin',), 'a_emax': ('a_emax',),
    'g_inv_t1': ('g_t1',), 'g_gain': ('g_gain',), 'g_vmin': ('g_vmin',), 'g_vmax': ('g_vmax',), 'g_r': ('g_t2', 'g_t3'),
    'g_omr': ('g_t2', 'g_t3'), 'g_inv_t3': ('g_t3',), 'g_dt': ('g_dt',),
    's_inv_tw': ('s_tw',), 's_kw': ('s_kw',), 's_r1': ('s_ll1_t1', 's_ll1_t2'), 's_omr1': ('s_ll1_t1', 's_ll1_t2'),
    's_inv_t3': ('s_ll1_t2',), 's_r2': ('s_ll2_t1', 's_ll2_t2'), 's_omr2': ('s_ll2_t1', 's_ll2_t2'),
    's_inv_t4': ('s_ll2_t2',), 's_hlim': ('s_hlim',),
}
# canonical state rows (include/dps.h enum dps_state)
X0_DEPS = (_BASE, _BASE, _BASE + ('x_d_t',), _BASE + ('x_q_t',), _BASE + ('x_d_st',), _BASE + ('x_q_st',),
           _EFD + ('a_k',), _EFD + ('a_emin', 'a_emax'), _PM + ('g_vmin', 'g_vmax'), _PM, _EFD, _EFD, _EFD)
AVR_SLOTS = ('a_vset', 'a_bias', 'a_r', 'a_omr', 'a_inv_tb', 'a_inv_te', 'a_k', 'a_emin', 'a_emax')
GOV_SLOTS = ('g_pref', 'g_inv_t1', 'g_gain', 'g_vmin', 'g_vmax', 'g_r', 'g_omr', 'g_inv_t3', 'g_dt')
PSS_SLOTS = ('s_inv_tw', 's_kw', 's_r1', 's_omr1', 's_inv_t3', 's_r2', 's_omr2', 's_inv_t4', 's_hlim')


def _bind(lib):
    if getattr(lib, '_prologue_bound', False):
        return
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    pi32 = ctypes.POINTER(i32)
    lib.dps_load_flow.argtypes = [i32, pi32, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, dbl, i32, pi32, vp]
    lib.dps_load_flow.restype = i32
    lib.dps_dynamic_inverse.argtypes = [i32, i64, vp, i64, i32, pi32, vp, i64, vp, i64, vp, vp, vp, vp]
    lib.dps_dynamic_inverse.restype = i32
    lib.dps_model_init.argtypes = [i32, i32, i64, pi32, pi32, vp, i64, vp, i64, vp, pi32, vp, vp, vp, vp, vp]
    lib.dps_model_init.restype = i32
    lib._prologue_bound = True


def _raw_attrs(g):
    """The DPS_N_RAW raw parameters of one generator in enum dps_raw order (None for absent controllers)."""
    v = [g.s_n, g.s_n_sys, g.x_d, g.x_q, g.x_d_t, g.x_q_t, g.x_d_st, g.x_q_st, g.t_d0_t, g.t_q0_t, g.t_d0_st, g.t_q0_st,
         g.h, g.d]
    if g.exciter:
        ll, pt = g.exciter.lead_lag, g.exciter.pt1_lim
        v += [ll.t_1, ll.t_2, pt.t_pt1, pt.gain_pt1, pt.lim_min, pt.lim_max, g.exciter.v_setpoint]
    else:
        v += [None] * 7
    if g.governor:
        pt, ll = g.governor.pt1_lim, g.governor.lead_lag
        v += [pt.t_pt1, pt.gain_pt1, pt.lim_min, pt.lim_max, ll.t_1, ll.t_2, g.governor.d_t]
    else:
        v += [None] * 7
    if g.stabilizer:
        st = g.stabilizer
        v += [st.washout.k_w, st.washout.t_w, st.lead_lag1.t_1, st.lead_lag1.t_2, st.lead_lag2.t_1, st.lead_lag2.t_2,
              st.limiter.limit]
    else:
        v += [None] * 7
    return v


def _cat_cols(cols):
    """[B,1] columns (B in {1, P}) -> [B,K]."""
    B = max(c.shape[0] for c in cols)
    return torch.cat([c.expand(B, 1) for c in cols], dim=1)


def _incidence(index, n, dev):
    m = torch.zeros((len(index), n), dtype=C128, device=dev)
    if index:
        m[torch.arange(len(index), device=dev), torch.tensor(index, device=dev)] = 1
    return m


class _Network(object):
    """Vectorised assembly of the admittance matrices and scheduled injections for the kernel path: the same terms as
    ``prologue._branch_terms`` / ``_assemble`` (simulator.py:362-384, 400-422), accumulated as a handful of small matrix
    products with 0/1 incidence matrices instead of one scatter per branch (deterministic, ~40 launches instead of ~1500;
    sums in a different order than the torch path: agreement ~1e-16 relative)."""

    def __init__(self, sim, dev):
        n = len(sim.busses)
        branches = list(sim.lines) + list(sim.trafos)
        self.F = _incidence([b.from_bus_id for b in branches], n, dev)
        self.T = _incidence([b.to_bus_id for b in branches], n, dev)
        nl = len(sim.lines)
        self.num = torch.tensor([1.0] * nl + [float(t.s_n) for t in sim.trafos], dtype=F64, device=dev)
        self.den = torch.tensor([1.0] * nl + [float(t.s_n_sys) for t in sim.trafos], dtype=F64, device=dev)
        self.shunts = [(bi, m) for bi, bus in enumerate(sim.busses) for m in bus.models if isinstance(m, Shunt)]
        self.gens = [(bi, m) for bi, bus in enumerate(sim.busses) for m in bus.models if isinstance(m, SynchMachine)]
        self.loads = [(bi, m) for bi, bus in enumerate(sim.busses) for m in bus.models if isinstance(m, Load)]
        self.inc_shunt = _incidence([bi for bi, _ in self.shunts], n, dev)
        self.inc_gen = _incidence([bi for bi, _ in self.gens], n, dev)
        self.inc_load = _incidence([bi for bi, _ in self.loads], n, dev)
        self.sys_gen = torch.tensor([float(m.s_n_sys) for _, m in self.gens], dtype=F64, device=dev)
        self.sys_load = torch.tensor([float(m.s_n_sys) for _, m in self.loads], dtype=F64, device=dev)
        self.key = (tuple((b.from_bus_id, b.to_bus_id) for b in branches), tuple(bi for bi, _ in self.shunts),
                    tuple(bi for bi, _ in self.gens), tuple(bi for bi, _ in self.loads))
        self.n, self.dev = n, dev

    def branches(self, sim, col):
        """[n,n] of lines and transformers (their parameters are read from the LAST trajectory: the reference's scatter with
        duplicate index tensors, static_models.py:321-322)."""
        br = list(sim.lines) + list(sim.trafos)
        if not br:
            return torch.zeros((self.n, self.n), dtype=C128, device=self.dev)
        r = torch.cat([col(b.r, F64, last_row=True) for b in br]).reshape(-1)
        x = torch.cat([col(b.x, F64, last_row=True) for b in br]).reshape(-1)
        bb = torch.cat([col(b.b, F64, last_row=True) for b in br]).reshape(-1)
        series = 1 / (r + 1j * x)
        y_off = (-series * self.num / self.den).unsqueeze(1)
        y_diag = ((series + 1j * bb / 2) * self.num / self.den).unsqueeze(1)
        Ft, Tt = self.F.t(), self.T.t()
        return Ft @ (y_off * self.T) + Tt @ (y_off * self.F) + Ft @ (y_diag * self.F) + Tt @ (y_diag * self.T)

    def static(self, yb, col):
        """Load-flow admittance [B,n,n]: branches + shunts."""
        if not self.shunts:
            return yb.unsqueeze(0)
        d = _cat_cols([col(m.y_shunt, C128) for _, m in self.shunts]) @ self.inc_shunt
        return yb.unsqueeze(0) + torch.diag_embed(d)

    def dynamic_base(self, yb, col):
        """Dynamic admittance without the loads [B,n,n]: + generator Norton admittances -j/x_d'' * S_n/S_sys
        (synchronous_machine.py:336)."""
        d = None
        if self.shunts:
            d = _cat_cols([col(m.y_shunt, C128) for _, m in self.shunts]) @ self.inc_shunt
        if self.gens:
            xdst = _cat_cols([col(m.x_d_st, C128) for _, m in self.gens])
            sn = _cat_cols([col(m.s_n, C128) for _, m in self.gens])
            g = (-1j / xdst * (sn / self.sys_gen)) @ self.inc_gen
            d = g if d is None else d + g
        return yb.unsqueeze(0) if d is None else yb.unsqueeze(0) + torch.diag_embed(d)

    def scheduled(self, col):
        """Scheduled bus injections [B,n] (static_models.py:165-176; synchronous_machine.py:347; static_models.py:495)."""
        s = torch.zeros((1, self.n), dtype=C128, device=self.dev)
        if self.gens:
            s = s + (_cat_cols([col(m.p_soll_mw, C128) for _, m in self.gens]) / self.sys_gen) @ self.inc_gen
        if self.loads:
            s = s + (-(self.load_power(col)) @ self.inc_load)
        return s

    def load_power(self, col):
        """Per-unit complex power of every load [B,L]."""
        p = _cat_cols([col(m.p_soll_mw, F64) for _, m in self.loads])
        q = _cat_cols([col(m.q_soll_mvar, F64) for _, m in self.loads])
        return (p + 1j * q) / self.sys_load


def _network_of(sim, cache, dev):
    net = cache.get('net_dev')
    key = (tuple((b.from_bus_id, b.to_bus_id) for b in list(sim.lines) + list(sim.trafos)),
           tuple(bi for bi, bus in enumerate(sim.busses) for m in bus.models if isinstance(m, Shunt)),
           tuple(bi for bi, bus in enumerate(sim.busses) for m in bus.models if isinstance(m, SynchMachine)),
           tuple(bi for bi, bus in enumerate(sim.busses) for m in bus.models if isinstance(m, Load)))
    if net is None or net.key != key or net.dev != dev:
        net = cache['net_dev'] = _Network(sim, dev)
    return net


def applicable(sim, device):
    """The kernel path runs on a CUDA device.  (Networks of more than MAX_BUS buses keep torch's batched LU for the load flow
    and the inverse -- their dense workspaces do not fit shared memory -- and take the model-initialisation kernel.)"""
    if not (str(device).startswith('cuda') and len(sim.busses) >= 1 and getattr(sim, 'prologue_kernels', True)):
        return False
    # dps_load_flow synchronises the stream once per Newton iteration: not inside a CUDA-graph capture
    return not torch.cuda.is_current_stream_capturing()


def _any_grad(values):
    return any(torch.is_tensor(v) and v.requires_grad for v in values)


class DeviceInitResult(prologue.InitResult):
    """``prologue.InitResult`` whose per-generator view lists (x0, slots, xst) are only built when somebody asks for them:
    ``Pss.prepare`` takes the kernel-layout arrays in ``packed`` directly."""
    __slots__ = ('packed', '_lazy', '_x0', '_slots', '_xst')

    def _build(self):
        if self._x0 is None:
            self._x0, self._slots, self._xst = self._lazy()

    x0 = property(lambda self: (self._build(), self._x0)[1], lambda self, v: setattr(self, '_x0', v))
    slots = property(lambda self: (self._build(), self._slots)[1], lambda self, v: setattr(self, '_slots', v))
    xst = property(lambda self: (self._build(), self._xst)[1], lambda self, v: setattr(self, '_xst', v))


def _classification(vary, base_vary, ctrl):
    """Which outputs differ per trajectory, from which raw parameters do: (per generator: rows of x0 that exist, batch-ness
    of each x0 row, slot names that exist, batch-ness of each slot)."""
    out = []
    for gi, flags in enumerate(ctrl):
        dep = {name: vary[gi * N_RAW + R[name]] for name in RAW}
        dep['base'] = base_vary
        rows = list(range(6)) + ([6, 7] if flags & CTRL_AVR else []) + ([8, 9] if flags & CTRL_GOV else []) + \
            ([10, 11, 12] if flags & CTRL_PSS else [])
        names = SLOTS[:15] + (AVR_SLOTS if flags & CTRL_AVR else ()) + (GOV_SLOTS if flags & CTRL_GOV else ()) + \
            (PSS_SLOTS if flags & CTRL_PSS else ())
        out.append((rows, [any(dep[k] for k in X0_DEPS[r]) for r in range(13)], names,
                    {name: any(dep[k] for k in SLOT_DEPS[name]) for name in names}))
    return out


def initialize(sim, device):
    """Kernel path of ``prologue.initialize``; returns None when an input requires grad (the caller then takes the
    differentiable torch path)."""
    lib = engine.load_library()
    _bind(lib)
    P = sim.parallel_sims
    n = len(sim.busses)
    gens = prologue.generators_of(sim)
    attrs = [a for _, g in gens for a in _raw_attrs(g)]
    net_in, dyn_in = prologue._network_inputs(sim), prologue._dynamic_inputs(sim)
    if torch.is_grad_enabled() and (_any_grad(net_in) or _any_grad(dyn_in) or _any_grad(attrs)):
        return None
    col = prologue._Cols(P, device, getattr(sim, '_uniform_cache', None))
    cache = sim.__dict__.setdefault('_prologue_cache', {})
    dev = torch.device(device)
    loads = [(bi, m) for bi, bus in enumerate(sim.busses) for m in bus.models if isinstance(m, Load)]
    res = DeviceInitResult()
    res._x0 = res._slots = res._xst = None

    with torch.no_grad(), torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

        def ptr(t):
            return ctypes.c_void_p(t.data_ptr()) if t is not None else None

        # ---- load flow -------------------------------------------------------------------------------------------
        lf_key, lf_keep, _ = prologue._fingerprint(net_in)
        hit = cache.get('lf_dev')
        if hit is not None and hit['key'] == lf_key:
            V, s_calc, iters = hit['out']
            lf_cached = True
        else:
            lf_cached = False
            net = _network_of(sim, cache, dev)
            yb = net.branches(sim, col)
            y_static = net.static(yb, col).contiguous()
            types = {'SL': 0, 'PV': 1, 'PQ': 2}
            bus_type = (ctypes.c_int32 * n)(*[types[b.lf_type] for b in sim.busses])
            s_soll = net.scheduled(col).contiguous()
            v_init = _cat_cols([col(b.voltage, C128) for b in sim.busses]).contiguous()
            Bs, Bv = s_soll.shape[0], v_init.shape[0]
            B = max(y_static.shape[0], Bs, Bv)
            if n <= MAX_BUS:
                V = torch.empty((B, n), dtype=C128, device=dev)
                s_calc = torch.empty((B, n), dtype=C128, device=dev)
                worst = torch.zeros(1, dtype=F64, device=dev)
                it = ctypes.c_int32(0)
                rc = lib.dps_load_flow(n, bus_type, B, ptr(y_static), y_static.shape[0], ptr(s_soll), Bs, ptr(v_init), Bv,
                                       ptr(V), ptr(s_calc), ptr(worst), 1e-8, 600, ctypes.byref(it), stream)
                if rc == 4:
                    raise ValueError('Load flow did not converge')
                engine._check(lib, rc, 'dps_load_flow')
                iters = int(it.value)
            else:                           # large network: torch's batched dense solve (same iteration, prologue.load_flow)
                V, s_calc, iters = prologue.load_flow(sim, col, y_static)
                V, s_calc = V.contiguous(), s_calc.contiguous()
            cache.pop('dyn_dev', None)
            cache['lf_dev'] = {'key': lf_key, 'keep': lf_keep, 'out': (V, s_calc, iters)}
        res.lf_iters, res.s_calc, res.bus_v = iters, s_calc, V

        # ---- loads -> admittances, Z0 = inv(Y_dyn) ----------------------------------------------------------------
        dyn_key, dyn_keep, _ = prologue._fingerprint(dyn_in)
        hit = cache.get('dyn_dev')
        if hit is not None and lf_cached and hit['key'] == dyn_key:
            y_load, res.Ydiag, res.Z0 = hit['out']
        else:
            net = _network_of(sim, cache, dev)
            y_base = net.dynamic_base(net.branches(sim, col), col).contiguous()
            nl = len(loads)
            if nl:
                s_load = net.load_power(col).contiguous()
                Bl = s_load.shape[0]
            else:
                Bl, s_load = 1, None
            Bd = max(y_base.shape[0], Bl, V.shape[0])
            load_bus = (ctypes.c_int32 * max(nl, 1))(*[bi for bi, _ in loads])
            y_load = torch.empty((Bd, max(nl, 1)), dtype=C128, device=dev)
            ydiag = torch.empty((Bd, n), dtype=C128, device=dev)
            z0 = torch.empty((Bd, n, n), dtype=C128, device=dev)
            if n <= MAX_BUS:
                rc = lib.dps_dynamic_inverse(n, Bd, ptr(y_base), y_base.shape[0], nl, load_bus, ptr(s_load), Bl, ptr(V),
                                             V.shape[0], ptr(y_load), ptr(ydiag), ptr(z0), stream)
                engine._check(lib, rc, 'dps_dynamic_inverse')
                y_load = y_load[:max(Bl, V.shape[0])]       # a load's admittance only depends on its own data and on V
            else:                           # large network: static_models.py:525-527 vectorised, torch.linalg.inv
                y_dyn = y_base
                if nl:
                    v_l = V[:, [bi for bi, _ in loads]]
                    y_load = 1 / torch.conj(torch.abs(v_l) ** 2 / s_load)
                    y_dyn = y_base + torch.diag_embed(y_load @ net.inc_load)
                ydiag = torch.diagonal(y_dyn, dim1=1, dim2=2).contiguous()
                z0 = torch.linalg.inv(y_dyn)
            res.Ydiag, res.Z0 = ydiag, z0
            cache['dyn_dev'] = {'key': dyn_key, 'keep': dyn_keep, 'out': (y_load, ydiag, z0)}
        for li, (_, m) in enumerate(loads):
            m.y_load = y_load[:, li:li + 1]

        # ---- model initialisation -----------------------------------------------------------------------------------
        G = len(gens)
        res.gens = [g for _, g in gens]
        res.gen_bus = [bi for bi, _ in gens]
        res.ctrl = [(CTRL_AVR if g.exciter else 0) | (CTRL_GOV if g.governor else 0) | (CTRL_PSS if g.stabilizer else 0)
                    for g in res.gens]
        # raw parameter table: per-attribute memo on identity + version (an identification loop replaces a handful of
        # per-trajectory tensors per run; the ~350 others keep their columns)
        zero1 = cache.get('zero1')
        if zero1 is None or zero1.device != dev:
            zero1 = cache['zero1'] = torch.zeros((1, 1), dtype=F64, device=dev)
        memo = cache.setdefault('raw_cols', {})
        cols, ukey = [], []
        for i, a in enumerate(attrs):
            k = (id(a), a._version) if torch.is_tensor(a) else a
            hit = memo.get(i)
            if hit is not None and hit[0] == k and (hit[2] is a or not torch.is_tensor(a)):
                c = hit[1]
            else:
                c = zero1 if a is None else col(a, F64)
                memo[i] = (k, c, a)
            cols.append(c)
            ukey.append(k if c.shape[0] == 1 else None)
        vary = tuple(c.shape[0] > 1 for c in cols)
        hit = cache.get('raw_u')
        if hit is not None and hit[0] == ukey:
            raw_u = hit[1]
        else:
            raw_u = torch.cat([zero1 if v else c for c, v in zip(cols, vary)]).reshape(G, N_RAW).contiguous()
            cache['raw_u'] = (ukey, raw_u)
        rows = [c.reshape(1, -1) for c, v in zip(cols, vary) if v]
        raw_v = torch.cat(rows, dim=0) if rows else None
        base_vary = V.shape[0] > 1 or s_calc.shape[0] > 1
        ckey = (vary, base_vary, tuple(res.ctrl), tuple(res.gen_bus))
        hit = cache.get('raw_class')
        if hit is None or hit[0] != ckey:
            idx, raw_vary = 0, []
            for v in vary:
                raw_vary.append(idx if v else -1)
                idx += 1 if v else 0
            cls = _classification(vary, base_vary, res.ctrl)
            B = P if (base_vary or any(vary)) else 1
            slot_vary, vrows = [], []
            for gi, (_, _, names, sv) in enumerate(cls):
                for k, name in enumerate(SLOTS):
                    if B > 1 and sv.get(name, False):
                        slot_vary.append(len(vrows))
                        vrows.append(gi * engine.N_SLOTS + k)
                    else:
                        slot_vary.append(-1)
            inc = torch.zeros((G, n), dtype=C128, device=dev)          # generator -> bus incidence
            inc[torch.arange(G, device=dev), torch.tensor(res.gen_bus, device=dev)] = 1
            hit = (ckey, {'a_bus': (ctypes.c_int32 * G)(*res.gen_bus), 'a_ctrl': (ctypes.c_int32 * G)(*res.ctrl),
                          'a_vary': (ctypes.c_int32 * (G * N_RAW))(*raw_vary), 'cls': cls, 'B': B, 'slot_vary': slot_vary,
                          'vrows': torch.tensor(vrows, dtype=torch.long, device=dev) if vrows else None, 'inc': inc})
            cache['raw_class'] = hit
        info = hit[1]
        B = info['B']
        x0 = torch.empty((G, engine.N_STATES, B), dtype=F64, device=dev)
        slots = torch.empty((G, engine.N_SLOTS, B), dtype=F64, device=dev)
        inj = torch.empty((G, 2, B), dtype=F64, device=dev)
        rc = lib.dps_model_init(G, n, B, info['a_bus'], info['a_ctrl'], ptr(V), V.shape[0], ptr(s_calc), s_calc.shape[0],
                                ptr(raw_u), info['a_vary'], ptr(raw_v), ptr(x0), ptr(slots), ptr(inj), stream)
        engine._check(lib, rc, 'dps_model_init')
        res.inj0 = torch.complex(inj[:, 0, :], inj[:, 1, :]).t()                  # [B, G] Norton currents at x0
        # what Pss.prepare packs for the stepping kernel, already in its layout
        res.packed = {'x0': x0, 'slots': slots, 'slot_vary': info['slot_vary'], 'vrows': info['vrows'],
                      'ctrl_key': tuple(k for gi in range(G) for k in ukey[gi * N_RAW + 14:(gi + 1) * N_RAW]),
                      'inj_bus': torch.matmul(res.inj0, info['inc'])}             # [B, n] injections per bus
        cls = info['cls']

        def views():
            """[B,1] views with the torch path's structure: B = P exactly for what depends on a per-trajectory input."""
            x0s, sls, xst = [], [], []
            for gi, g in enumerate(res.gens):
                rows, xv, names, sv = cls[gi]
                x0s.append([x0[gi, k, :(B if xv[k] else 1)].reshape(-1, 1) for k in rows])
                sls.append({name: slots[gi, SLOT_INDEX[name], :(B if sv[name] else 1)].reshape(-1, 1) for name in names})
                xst.append((col(g.x_d_st, C128), col(g.x_q_st, C128)))
            return x0s, sls, xst
        res._lazy = views

        # expose initial values on the containers, as the reference does (complex, like its own attributes)
        xc = x0[:, :6, :].to(C128)                                                 # machine states
        exposed = ('pm0', 'efd0', 'a_bias', 'g_pref')
        sc = slots[:, [SLOT_INDEX[k] for k in exposed], :].to(C128)
        for gi, g in enumerate(res.gens):
            _, xv, _, sv = cls[gi]
            cx = [xc[gi, k, :(B if xv[k] else 1)].reshape(-1, 1) for k in range(6)]
            g.omega, g.delta, g.e_q_t, g.e_d_t, g.e_q_st, g.e_d_st = cx
            g.p_m = sc[gi, 0, :(B if sv['pm0'] else 1)].reshape(-1, 1)
            g.e_fd = sc[gi, 1, :(B if sv['efd0'] else 1)].reshape(-1, 1)
            if g.exciter:
                g.exciter.bias = sc[gi, 2, :(B if sv['a_bias'] else 1)].reshape(-1, 1)
            if g.governor:
                g.governor.p_ref = sc[gi, 3, :(B if sv['g_pref'] else 1)].reshape(-1, 1)
    return res
