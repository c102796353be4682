"""Recorder support: turning the user's ``record_func(sim)`` into kernel recorder channels.

Reference contract (simulator.py:476-483, 540-546): ``record_func`` is an arbitrary Python callable that is
invoked after *every* step, reads attributes of the live model objects
(``sim.busses[1].models[0].omega.real``, ``sim.busses[2].voltage.imag``, ``...exciter.pt1_lim.state_1`` ...) and
returns a list of ``[P,1]`` tensors; ``run()`` stacks them to ``[batch, timestep, value]``.

A persistent kernel cannot call back into Python 2000 times, so the function is *traced once*: it is called on a
view of the simulation whose state / algebraic attributes are symbolic :class:`Expr` leaves.  The set of leaves it
touched becomes the kernel's recorder channel list (coalesced ``rec[channel][step][sim]`` stores); any arithmetic the
function did on top of the leaves (``abs``, ``+``, ``*`` with parameters, ``torch.*`` calls) is replayed afterwards
with torch on whole ``[T,P,1]`` channel arrays, which is equivalent because every such op is element-wise.  Anything
the tracer cannot express raises -- there is no per-step Python fallback.
"""
import torch

from . import engine
from .components import SynchMachine

# field ids: must match enum dps_state / dps_field in include/dps.h
STATE_FIELDS = {'omega': 0, 'delta': 1, 'e_q_t': 2, 'e_d_t': 3, 'e_q_st': 4, 'e_d_st': 5}
ALG_FIELDS = {'p_e': 13, 'i_d': 14, 'i_q': 15, 'v_re': 16, 'v_im': 17, 'e_fd': 18, 'p_m': 19}
BLOCK_FIELDS = {('exciter', 'lead_lag'): 6, ('exciter', 'pt1_lim'): 7,
                ('governor', 'pt1_lim'): 8, ('governor', 'lead_lag'): 9,
                ('stabilizer', 'washout'): 10, ('stabilizer', 'lead_lag1'): 11, ('stabilizer', 'lead_lag2'): 12}
NAMED_FIELDS = dict(STATE_FIELDS, **ALG_FIELDS)
NAMED_FIELDS.update({'avr_ll': 6, 'avr_pt1': 7, 'gov_pt1': 8, 'gov_ll': 9, 'pss_w': 10, 'pss_ll1': 11, 'pss_ll2': 12})


class Expr(object):
    """Symbolic value: a leaf (recorder channel) or a deferred torch operation on other values."""
    __slots__ = ('op', 'args', 'kwargs', 'leaf')

    def __init__(self, op=None, args=(), kwargs=None, leaf=None):
        self.op, self.args, self.kwargs, self.leaf = op, args, kwargs or {}, leaf

    # ---- leaves ------------------------------------------------------------------------------------------------
    @staticmethod
    def real_leaf(key):
        return Expr(leaf=('real', key))

    @staticmethod
    def complex_leaf(key_re, key_im):
        return Expr(leaf=('complex', key_re, key_im))

    @property
    def real(self):
        if self.leaf:
            return self if self.leaf[0] == 'real' else Expr.real_leaf(self.leaf[1])
        return Expr(_attr_real, (self,))

    @property
    def imag(self):
        if self.leaf:
            if self.leaf[0] == 'real':
                return Expr(torch.zeros_like, (self,))     # states carry an identically-zero imaginary part
            return Expr.real_leaf(self.leaf[2])
        return Expr(_attr_imag, (self,))

    # ---- deferred arithmetic -----------------------------------------------------------------------------------
    def _bin(self, other, fn, swap=False):
        return Expr(fn, (other, self) if swap else (self, other))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, _radd, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, _rsub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, _rmul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, _rdiv, True)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __neg__(self): return Expr(torch.neg, (self,))
    def __abs__(self): return Expr(torch.abs, (self,))

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        return Expr(func, tuple(args), kwargs)

    # ---- evaluation --------------------------------------------------------------------------------------------
    def leaves(self, out):
        if self.leaf:
            out.extend(self.leaf[1:])
            return
        for a in list(self.args) + list(self.kwargs.values()):
            _collect(a, out)

    def evaluate(self, fetch):
        if self.leaf:
            if self.leaf[0] == 'real':
                return fetch(self.leaf[1])
            return torch.complex(fetch(self.leaf[1]), fetch(self.leaf[2]))
        args = [_eval(a, fetch) for a in self.args]
        kwargs = {k: _eval(v, fetch) for k, v in self.kwargs.items()}
        return self.op(*args, **kwargs)


def _attr_real(t): return t.real
def _attr_imag(t): return t.imag if t.is_complex() else torch.zeros_like(t)
def _radd(a, b): return a + b
def _rsub(a, b): return a - b
def _rmul(a, b): return a * b
def _rdiv(a, b): return a / b


def _collect(a, out):
    if isinstance(a, Expr):
        a.leaves(out)
    elif isinstance(a, (list, tuple)):
        for v in a:
            _collect(v, out)


def _eval(a, fetch):
    if isinstance(a, Expr):
        return a.evaluate(fetch)
    if isinstance(a, (list, tuple)):
        return type(a)(_eval(v, fetch) for v in a)
    return a


# ------------------------------------------------------------------------------------------------------------------
# views handed to the user's record function
# ------------------------------------------------------------------------------------------------------------------
class _BlockView(object):
    def __init__(self, block, key):
        object.__setattr__(self, '_block', block)
        object.__setattr__(self, '_key', key)

    def __getattr__(self, name):
        if name == 'state_1' and self._key is not None:
            return Expr.real_leaf(self._key)
        return getattr(self._block, name)


class _ControllerView(object):
    def __init__(self, ctrl, gen_index, role):
        object.__setattr__(self, '_ctrl', ctrl)
        object.__setattr__(self, '_gi', gen_index)
        object.__setattr__(self, '_role', role)

    def __getattr__(self, name):
        fid = BLOCK_FIELDS.get((self._role, name))
        val = getattr(self._ctrl, name)
        if fid is not None:
            return _BlockView(val, ('gen', self._gi, fid))
        return val

    def __bool__(self):
        return True


class _MachineView(object):
    def __init__(self, machine, gen_index):
        object.__setattr__(self, '_m', machine)
        object.__setattr__(self, '_gi', gen_index)

    def __getattr__(self, name):
        gi = self._gi
        if name in STATE_FIELDS:
            return Expr.real_leaf(('gen', gi, STATE_FIELDS[name]))
        if name in ('p_e', 'i_d', 'i_q', 'e_fd', 'p_m'):
            return Expr.real_leaf(('gen', gi, ALG_FIELDS[name]))
        if name == 'v_bb':
            return Expr.complex_leaf(('gen', gi, ALG_FIELDS['v_re']), ('gen', gi, ALG_FIELDS['v_im']))
        val = getattr(self._m, name)
        if name in ('exciter', 'governor', 'stabilizer') and val is not None:
            return _ControllerView(val, gi, name)
        return val


class _BusView(object):
    def __init__(self, bus, bus_index, gen_index_of):
        object.__setattr__(self, '_bus', bus)
        object.__setattr__(self, '_bi', bus_index)
        object.__setattr__(self, '_gmap', gen_index_of)

    @property
    def models(self):
        return [_MachineView(m, self._gmap[id(m)]) if isinstance(m, SynchMachine) else m for m in self._bus.models]

    @property
    def voltage(self):
        return Expr.complex_leaf(('bus', self._bi, 0), ('bus', self._bi, 1))

    def __getattr__(self, name):
        return getattr(self._bus, name)


class SimView(object):
    """What ``record_func`` sees instead of the simulation while it is being traced."""

    def __init__(self, sim, gens):
        object.__setattr__(self, '_sim', sim)
        gmap = {id(g): i for i, g in enumerate(gens)}
        object.__setattr__(self, 'busses', [_BusView(b, i, gmap) for i, b in enumerate(sim.busses)])

    def __getattr__(self, name):
        return getattr(self._sim, name)


# ------------------------------------------------------------------------------------------------------------------
class RecordPlan(object):
    """Result of tracing: the kernel channel list and how to build the user-visible result from it."""

    def __init__(self, outputs, channel_keys):
        self.outputs = outputs              # list of Expr (one per recorded value)
        self.channel_keys = channel_keys    # ordered unique leaf keys: ('gen', g, field) | ('bus', bus, 0|1)
        self.direct = all(o.leaf and o.leaf[0] == 'real' for o in outputs) and \
            [o.leaf[1] for o in outputs] == channel_keys

    def assemble(self, rec):
        """rec: [n_chan, T, P] -> [P, T, n_values] (reference layout, simulator.py:546)."""
        if self.direct:
            return rec.permute(2, 1, 0)
        index = {k: i for i, k in enumerate(self.channel_keys)}

        def fetch(key):
            return rec[index[key]].unsqueeze(-1)            # [T, P, 1]

        vals = []
        for o in self.outputs:
            v = o.evaluate(fetch)
            if v.dim() == 2:
                v = v.unsqueeze(-1)
            vals.append(v.expand(rec.shape[1], rec.shape[2], 1))
        return torch.stack(vals, dim=0).squeeze(-1).permute(2, 1, 0)


def trace(sim, gens, record_func):
    view = SimView(sim, gens)
    try:
        outputs = record_func(view)
    except Exception as exc:  # noqa
        raise engine.EngineError(
            "backend='cuda' could not trace the record function (%s: %s). It may read machine states/algebraics, "
            "controller block states and bus voltages and combine them with element-wise torch ops."
            % (type(exc).__name__, exc))
    if outputs is None:
        raise TypeError('record function returned None')
    outs = []
    for o in outputs:
        if not isinstance(o, Expr):
            const = o
            o = Expr(_constant, (const,))
        outs.append(o)
    return _finish(outs)


def _constant(c):
    return c


def from_names(sim, gens, names):
    """Channels by name: '<generator name>.<field>' or 'bus:<bus name>.v_re|v_im' (test / bench convenience)."""
    gidx = {g.name: i for i, g in enumerate(gens)}
    outs = []
    for n in names:
        if n.startswith('bus:'):
            bname, f = n[4:].rsplit('.', 1)
            outs.append(Expr.real_leaf(('bus', sim.bus_idxs[bname], {'v_re': 0, 'v_im': 1}[f])))
        else:
            gname, f = n.rsplit('.', 1)
            outs.append(Expr.real_leaf(('gen', gidx[gname], NAMED_FIELDS[f])))
    return _finish(outs)


def _finish(outs):
    keys = []
    for o in outs:
        ks = []
        o.leaves(ks)
        for k in ks:
            if k not in keys:
                keys.append(k)
    return RecordPlan(outs, keys)
