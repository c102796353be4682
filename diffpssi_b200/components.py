"""Grid components: the parameter containers of the ``Pss`` API.

These mirror the *attribute surface* of the reference model classes, because user code pokes them
directly (``sim.busses[1].models[0].h = torch.tensor(..., requires_grad=True)``,
``...stabilizer.washout.k_w = ...`` -- reference examples/models/ibb_model/ibb_opt.py:48-57 and
examples/models/k2a/k2a_opt.py:111-185) and record functions read them
(``sim.busses[0].models[0].omega.real``).  Unlike the reference classes they carry **no equations**:
the differential equations, current injections and block outputs of

    synchronous_machine.py:138-277, blocks.py:103-436, exciters.py:74-115,
    governors.py:77-123, stabilizers.py:75-121, static_models.py:141-163

are evaluated by the CUDA step kernel (``csrc/dps_kernels.cu``); the one-off initialisation formulas
live in ``prologue.py``.  Containers only (i) take the same constructor arguments / dict keys,
(ii) broadcast scalars to ``[parallel_sims, 1]`` tensors with the reference's dtypes
(complex128 for machine quantities, float64 for controller blocks) and (iii) expose the state /
algebraic attributes, which the simulator fills after ``initialize()`` and after ``run()``.
"""
import weakref

import torch

C128 = torch.complex128
F64 = torch.float64


# Tensors this package itself created by broadcasting ONE number over the batch: known to be batch-uniform without looking
# at their contents (the prologue otherwise pays a device comparison + synchronisation per parameter tensor on its first
# call: ~650 of them for the IEEE 39-bus grid).  Keyed by identity, valid for the version counter recorded here.
_KNOWN_UNIFORM = {}


def mark_uniform(t):
    k = id(t)
    _KNOWN_UNIFORM[k] = (weakref.ref(t, lambda _r, k=k: _KNOWN_UNIFORM.pop(k, None)), t._version)
    return t


def known_uniform(t):
    hit = _KNOWN_UNIFORM.get(id(t))
    return hit is not None and hit[0]() is t and hit[1] == t._version


def _bcast(value, P, dtype, device):
    """Scalar (or tensor) -> [P,1] tensor.  A user-supplied tensor is kept as is (it may be an autograd leaf)."""
    if torch.is_tensor(value):
        return value
    return mark_uniform(torch.full((P, 1), value, dtype=dtype, device=device))


class _Component(object):
    """Base: declarative parameter tables instead of hand-written constructors."""
    DICT_KEYS = ()          # ((dict key, attribute name, kwarg name), ...)
    PARALLEL = ()           # attributes broadcast by enable_parallel_simulation
    DTYPE = F64

    def _take(self, param_dict, kwargs):
        if param_dict is None:
            param_dict = {dk: kwargs.get(kw) for dk, _attr, kw in self.DICT_KEYS}
        for dk, attr, _kw in self.DICT_KEYS:
            setattr(self, attr, param_dict[dk] if dk in param_dict else None)
        return param_dict

    def enable_parallel_simulation(self, parallel_sims, device='cpu'):
        self.parallel_sims = parallel_sims
        for attr in self.PARALLEL:
            setattr(self, attr, _bcast(getattr(self, attr), parallel_sims, self.DTYPE, device))


# --------------------------------------------------------------------------------------------------
# events   (reference: static_models.py:10-95)
# --------------------------------------------------------------------------------------------------
class ScEvent(object):
    """Three-phase short circuit at ``bus`` (index).  Active iff ``start < t <= end`` evaluated on the
    float time grid ``np.arange(0, sim_time, dt)`` -- reference static_models.py:36-49."""

    def __init__(self, start_time, end_time, bus):
        self.start_time, self.end_time, self.bus = start_time, end_time, bus

    def is_active(self, t):
        return bool(self.start_time < t <= self.end_time)


class ParamEvent(object):
    """``setattr(model, name, value)`` once when ``t >= start`` -- reference static_models.py:52-95.
    The simulator splits the run into kernel launches at these steps."""

    def __init__(self, start_time, model, param_name, value):
        self.start_time, self.model, self.param_name, self.value = start_time, model, param_name, value
        self.handled = False


# --------------------------------------------------------------------------------------------------
# static network   (reference: static_models.py:100-694)
# --------------------------------------------------------------------------------------------------
class Bus(_Component):
    DICT_KEYS = (('name', 'name', 'name'), ('V_n', 'v_n', 'v_n'))
    PARALLEL = ('v_n', 'voltage')
    DTYPE = C128

    def __init__(self, param_dict=None, **kw):
        self._take(param_dict, kw)
        self.parallel_sims = None
        self.models = []
        self.diff_models = []
        self.lf_type = 'PQ'
        self.voltage = 1.0

    def add_model(self, model):
        self.models.append(model)
        if isinstance(model, SynchMachine):
            self.diff_models.append(model)


class Line(_Component):
    """pi-line.  Unit conversion as reference static_models.py:286-294: 'Ohm' values are divided by the
    system base impedance, 'p.u.' values are re-based from the line's own (S_n, V_n)."""
    PARALLEL = ('r', 'x', 'b')

    def __init__(self, s_n_sys, v_n_sys, param_dict=None, **kw):
        if param_dict is None:
            param_dict = {'name': kw.get('name'), 'from_bus': kw.get('from_bus'), 'to_bus': kw.get('to_bus'),
                          'length': kw.get('length'), 'S_n': kw.get('s_n'), 'V_n': kw.get('v_n'),
                          'unit': kw.get('unit'), 'R': kw.get('r'), 'X': kw.get('x'), 'B': kw.get('b')}
        self.name = param_dict['name']
        self.from_bus_name, self.to_bus_name = param_dict['from_bus'], param_dict['to_bus']
        self.from_bus_id = self.to_bus_id = None
        length, unit = param_dict['length'], param_dict['unit']
        s_n = param_dict.get('S_n', s_n_sys)
        v_n = param_dict.get('V_n', v_n_sys)
        z_sys = v_n_sys ** 2 / s_n_sys
        if unit == 'Ohm':
            self.r = param_dict['R'] * length / z_sys
            self.x = param_dict['X'] * length / z_sys
            self.b = param_dict['B'] * length * z_sys
        elif unit == 'p.u.':
            z_own = v_n ** 2 / s_n
            self.r = param_dict['R'] * z_own / z_sys * length
            self.x = param_dict['X'] * z_own / z_sys * length
            self.b = param_dict['B'] / z_own * z_sys * length
        else:
            raise ValueError('Unit not supported')


class Transformer(_Component):
    """Series impedance on its own S_n base, no tap -- reference static_models.py:328-427."""
    DICT_KEYS = (('name', 'name', 'name'), ('from_bus', 'from_bus_name', 'from_bus'),
                 ('to_bus', 'to_bus_name', 'to_bus'), ('S_n', 's_n', 's_n'), ('R', 'r', 'r'), ('X', 'x', 'x'),
                 ('V_n_from', 'v_n_from', 'v_n_from'), ('V_n_to', 'v_n_to', 'v_n_to'))
    PARALLEL = ('r', 'x', 'v_n_from', 'v_n_to')

    def __init__(self, s_n_sys, param_dict=None, **kw):
        pd = self._take(param_dict, kw)
        b = pd.get('B', kw.get('b', 0))
        self.b = 0 if b is None else b
        self.from_bus_id = self.to_bus_id = None
        self.s_n_sys = s_n_sys


class Load(_Component):
    """Constant-impedance load; its admittance is fixed from the load-flow voltage
    (reference static_models.py:515-527, evaluated in prologue.py)."""
    DICT_KEYS = (('name', 'name', 'name'), ('bus', 'bus', 'bus'), ('P', 'p_soll_mw', 'p'),
                 ('Q', 'q_soll_mvar', 'q'), ('model', 'model', 'model'))
    PARALLEL = ('p_soll_mw', 'q_soll_mvar')

    def __init__(self, s_n_sys, param_dict=None, **kw):
        self._take(param_dict, kw)
        if not self.model == 'Z':
            raise ValueError('Only Z model is supported for loads')
        self.s_n_sys = s_n_sys
        self.y_load = None


class Shunt(_Component):
    DICT_KEYS = (('name', 'name', 'name'), ('bus', 'bus', 'bus'), ('V_n', 'v_n', 'v_n'),
                 ('Q', 'q_soll_mvar', 'q'), ('model', 'model', 'model'))
    PARALLEL = ('q_soll_mvar',)

    def __init__(self, s_n_sys, param_dict=None, **kw):
        self._take(param_dict, kw)
        if not self.model == 'Z':
            raise ValueError('Only Z model is supported for shunts')
        self.s_n_sys = s_n_sys
        # Parity quirk, inherited on purpose: the reference's torch path builds the shunt power with
        # ``torch.tensor(python_complex)`` (static_models.py:617-619), i.e. in *complex64*, and only then
        # widens to complex128 -- y_shunt carries float32 rounding (0.2222222238779068 for 200 Mvar / 900 MVA).
        # Parity is judged against that path, so the same rounding is applied here.
        s_shunt = torch.tensor(-1j * self.q_soll_mvar / self.s_n_sys)
        self.y_shunt = 1 / torch.conj(1 / s_shunt)

    def enable_parallel_simulation(self, parallel_sims, device='cpu'):
        super().enable_parallel_simulation(parallel_sims, device)
        self.y_shunt = mark_uniform(torch.ones((parallel_sims, 1), dtype=C128, device=device) * self.y_shunt.to(device))


# --------------------------------------------------------------------------------------------------
# controller blocks   (reference: blocks.py:103-436)  -- one state each, float64 parameters
# --------------------------------------------------------------------------------------------------
class PIController(_Component):
    """Proportional-integral block, ``output = k_p * input + k_i * state`` (reference blocks.py:7-100).  The reference
    defines it for inverter controls but ships no model that uses it; it is a parameter container here like the other
    blocks, so that a user-defined model can carry it."""
    PARALLEL = ('k_p', 'k_i', 'state_1')

    def __init__(self, k_p, k_i):
        self.k_p, self.k_i = k_p, k_i
        self.input = 0
        self.state_1 = 0


class PT1Limited(_Component):
    PARALLEL = ('t_pt1', 'gain_pt1', 'lim_min', 'lim_max')

    def __init__(self, t_pt1, gain_pt1, lim_min, lim_max):
        self.t_pt1, self.gain_pt1, self.lim_min, self.lim_max = t_pt1, gain_pt1, lim_min, lim_max
        self.state_1 = 0


class LeadLag(_Component):
    PARALLEL = ('t_1', 't_2')

    def __init__(self, t_1, t_2):
        self.t_1, self.t_2 = t_1, t_2
        self.state_1 = 0


class Washout(_Component):
    PARALLEL = ('k_w', 't_w')

    def __init__(self, k_w, t_w):
        self.k_w, self.t_w = k_w, t_w
        self.state_1 = 0


class Limiter(_Component):
    PARALLEL = ('limit',)

    def __init__(self, limit):
        self.limit = limit


class _Controller(_Component):
    BLOCKS = ()

    def enable_parallel_simulation(self, parallel_sims, device='cpu'):
        super().enable_parallel_simulation(parallel_sims, device)
        for b in self.BLOCKS:
            getattr(self, b).enable_parallel_simulation(parallel_sims, device)


class SEXS(_Controller):
    """Simplified excitation system: LeadLag(T_a, T_b) -> PT1Limited(T_e, K, E_min, E_max); state order
    [lead_lag, pt1_lim] (reference exciters.py:71-100)."""
    DICT_KEYS = (('name', 'name', 'name'), ('gen', 'gen', 'gen'), ('T_a', 't_a', 't_a'), ('T_b', 't_b', 't_b'),
                 ('T_e', 't_e', 't_e'), ('K', 'k', 'k'), ('E_min', 'e_min', 'e_min'), ('E_max', 'e_max', 'e_max'))
    PARALLEL = ('v_setpoint', 'bias')
    BLOCKS = ('lead_lag', 'pt1_lim')

    def __init__(self, param_dict=None, **kw):
        self._take(param_dict, kw)
        self.lead_lag = LeadLag(t_1=self.t_a, t_2=self.t_b)
        self.pt1_lim = PT1Limited(t_pt1=self.t_e, gain_pt1=self.k, lim_min=self.e_min, lim_max=self.e_max)
        self.v_setpoint = 0.0
        self.bias = 0.0


class TGOV1(_Controller):
    """Steam governor: PT1Limited(T_1, 1/R, V_min, V_max) -> LeadLag(T_2, T_3), minus D_t*omega; state order
    [pt1_lim, lead_lag].  ``1/R`` is fixed at construction like the reference (governors.py:71-72)."""
    DICT_KEYS = (('name', 'name', 'name'), ('gen', 'gen', 'gen'), ('R', 'r', 'r'), ('D_t', 'd_t', 'd_t'),
                 ('T_1', 't_1', 't_1'), ('T_2', 't_2', 't_2'), ('T_3', 't_3', 't_3'),
                 ('V_min', 'v_min', 'v_min'), ('V_max', 'v_max', 'v_max'))
    PARALLEL = ('p_ref',)
    BLOCKS = ('pt1_lim', 'lead_lag')

    def __init__(self, param_dict=None, **kw):
        self._take(param_dict, kw)
        self.pt1_lim = PT1Limited(t_pt1=self.t_1, gain_pt1=1 / self.r, lim_min=self.v_min, lim_max=self.v_max)
        self.lead_lag = LeadLag(t_1=self.t_2, t_2=self.t_3)
        self.p_ref = 0.0


class STAB1(_Controller):
    """PSS: Washout(K, T) -> LeadLag(T_1, T_3) -> LeadLag(T_2, T_4) -> Limiter(H_lim)
    (block wiring as reference stabilizers.py:70-73)."""
    DICT_KEYS = (('name', 'name', 'name'), ('gen', 'gen', 'gen'), ('K', 'k_w', 'k'), ('T', 't_w', 't'),
                 ('T_1', 't_1', 't_1'), ('T_2', 't_2', 't_2'), ('T_3', 't_3', 't_3'), ('T_4', 't_4', 't_4'),
                 ('H_lim', 'h_lim', 'h_lim'))
    BLOCKS = ('washout', 'lead_lag1', 'lead_lag2', 'limiter')

    def __init__(self, param_dict=None, **kw):
        self._take(param_dict, kw)
        self.washout = Washout(k_w=self.k_w, t_w=self.t_w)
        self.lead_lag1 = LeadLag(t_1=self.t_1, t_2=self.t_3)
        self.lead_lag2 = LeadLag(t_1=self.t_2, t_2=self.t_4)
        self.limiter = Limiter(limit=self.h_lim)


# --------------------------------------------------------------------------------------------------
# synchronous machine   (reference: synchronous_machine.py:8-388)
# --------------------------------------------------------------------------------------------------
class SynchMachine(_Component):
    DICT_KEYS = (('name', 'name', 'name'), ('bus', 'bus', 'bus'), ('S_n', 's_n', 's_n'), ('V_n', 'v_n', 'v_n'),
                 ('P', 'p_soll_mw', 'p'), ('V', 'v_soll', 'v'), ('H', 'h', 'h'), ('D', 'd', 'd'),
                 ('X_d', 'x_d', 'x_d'), ('X_q', 'x_q', 'x_q'), ('X_d_t', 'x_d_t', 'x_d_t'),
                 ('X_q_t', 'x_q_t', 'x_q_t'), ('X_d_st', 'x_d_st', 'x_d_st'), ('X_q_st', 'x_q_st', 'x_q_st'),
                 ('T_d0_t', 't_d0_t', 't_d0_t'), ('T_q0_t', 't_q0_t', 't_q0_t'),
                 ('T_d0_st', 't_d0_st', 't_d0_st'), ('T_q0_st', 't_q0_st', 't_q0_st'))
    STATES = ('omega', 'delta', 'e_q_t', 'e_d_t', 'e_q_st', 'e_d_st')
    ALGEBRAICS = ('p_m', 'p_e', 'e_fd', 'i_d', 'i_q', 'v_bb')
    PARALLEL = ('s_n', 'v_n', 'p_soll_mw', 'v_soll', 'h', 'd', 'x_d', 'x_q', 'x_d_t', 'x_q_t', 'x_d_st', 'x_q_st',
                't_d0_t', 't_q0_t', 't_d0_st', 't_q0_st') + STATES + ALGEBRAICS
    DTYPE = C128

    def __init__(self, f_n_sys, s_n_sys, v_n_sys, param_dict=None, **kw):
        self._take(param_dict, kw)
        self.parallel_sims = None
        self.fn, self.s_n_sys, self.v_n_sys = f_n_sys, s_n_sys, v_n_sys
        for a in self.STATES + self.ALGEBRAICS:
            setattr(self, a, 0)
        self.exciter = self.governor = self.stabilizer = None

    def add_exciter(self, exciter_model):
        self.exciter = exciter_model

    def add_governor(self, governor_model):
        self.governor = governor_model

    def add_pss(self, pss_model):
        self.stabilizer = pss_model

    def n_states(self):
        return 6 + (2 if self.exciter else 0) + (2 if self.governor else 0) + (3 if self.stabilizer else 0)
