"""``PowerSystemSimulation`` -- the reference's ``Pss`` API with ``backend='cuda'``.

Drop-in mirror of src/diffpssi/power_sim_lib/simulator.py:21-552 for the path this package accelerates:

    Pss(time_step, sim_time, parallel_sims, solver, grid_data=None, verbose=True, backend='cuda')
    add_bus / add_generator / add_load / add_shunt / add_line / add_transformer / add_exciter / add_governor /
    add_pss, set_slack_bus, add_sc_event, add_param_event, set_record_function, reset, initialize, run

``run()`` returns ``(time, result)`` with ``result[batch, timestep, value]`` exactly like simulator.py:546.
Everything before the time loop is host torch code (``prologue.py``); the loop itself is ONE call into the CUDA
engine (``engine.py`` -> ``libdps_b200.so``).  There is no numpy/torch stepping path in this package: the
reference's own backends are what ``backend='numpy'|'torch'`` mean, and asking for them here raises.
"""
import ctypes
import os
import time as _time

import numpy as np
import torch

from . import adjoint, engine, planning, prologue, prologue_device, recording
from .autograd import StepFunction
from .components import (Bus, Line, Transformer, Load, Shunt, SynchMachine, SEXS, TGOV1, STAB1, ScEvent,
                         ParamEvent, mark_uniform)

SOLVERS = ('euler', 'heun', 'rk4')            # solvers.py:226-230
BACKEND = os.environ.get('DIFFPSSI_FORCE_SIM_BACKEND', 'cuda')      # backend.py:11-14


def _tabular(value):
    """List-of-lists table (first row = keys) -> list of dicts (simulator.py:139-154)."""
    keys = value[0]
    return [dict(zip(keys, row)) for row in value[1:]]


class PowerSystemSimulation(object):

    def __init__(self, time_step, sim_time, parallel_sims, solver, grid_data=None, verbose=True,
                 backend=None, device=None, ckpt_stride=None, prologue_device=None):
        backend = backend or BACKEND
        if backend != 'cuda':
            raise ValueError("diffpssi_b200 implements backend='cuda' only (got %r); the numpy/torch backends are "
                             "the reference's own" % (backend,))
        self.backend = backend
        if device is None:
            device = 'cuda' if torch.cuda.is_available() else 'cpu'      # cpu: host-side logic only (tests)
        self.device = str(device)
        # Extension: where the once-per-run prologue (load flow, model initialisation, inv(Y)) is evaluated.  Default: the
        # run device.  'cpu' keeps the ~10^3 tiny tensor ops of the prologue off the GPU (no launch latency -- the better
        # choice for small batches) and uploads the packed kernel inputs; autograd still reaches the user's leaves.
        self.prologue_device = str(prologue_device) if prologue_device is not None else self.device
        self.time = np.arange(0, sim_time, time_step)                       # simulator.py:78
        self.time_step = time_step
        self.busses, self.non_slack_busses, self.bus_idxs = [], [], {}
        self.lines, self.trafos = [], []
        self.inverse_dynamic_y_matrix = None
        self.static_y_matrix = None
        self.sc_events, self.param_events = [], []
        self.parallel_sims = parallel_sims
        self.record_func = None
        self.record_channels = None
        self.verbose = verbose
        self.ckpt_stride = ckpt_stride
        forced = os.environ.get('DIFFPSSI_FORCE_INTEGRATOR')                # simulator.py:96-100
        if forced is not None:
            print('WARNING: FORCING THE USE OF THE {} INTEGRATOR. '
                  'THIS SHOULD ONLY HAPPEN FOR UNITTESTS'.format(forced))
            solver = forced
        if solver not in SOLVERS:
            raise KeyError(solver)
        self.solver = solver
        self.base_voltage = self.base_mva = self.fn = None
        self.last_init = None
        self.last_plan = None
        self._uniform_cache = {}
        self.timings = {}
        if grid_data is not None:
            self.create_grid(grid_data)

    # ------------------------------------------------------------------------------------------------------------
    # grid construction (simulator.py:113-348)
    # ------------------------------------------------------------------------------------------------------------
    def get_generator_by_name(self, name):
        for bus in self.busses:
            for model in bus.models:
                if getattr(model, 'name', None) == name:
                    return model
        return None

    def create_grid(self, grid_data):
        self.fn = grid_data['f']
        self.base_mva = grid_data['base_mva']
        self.base_voltage = grid_data['base_voltage']
        data = {}
        for key, value in grid_data.items():
            if isinstance(value, list) and all(isinstance(item, list) for item in value):
                data[key] = _tabular(value)
            elif isinstance(value, dict):
                data[key] = {sub: _tabular(tab) for sub, tab in value.items()}
            else:
                data[key] = value
        for d in data['busses']:
            self.add_bus(Bus(param_dict=d))
        gens = data.get('generators', [])
        if isinstance(gens, dict):
            for d in gens['GEN']:
                self.add_generator(SynchMachine(param_dict=d, s_n_sys=self.base_mva, v_n_sys=self.base_voltage,
                                                f_n_sys=self.fn))
        for d in data.get('loads', []):
            self.add_load(Load(param_dict=d, s_n_sys=self.base_mva))
        for d in data.get('shunts', []):
            self.add_shunt(Shunt(param_dict=d, s_n_sys=self.base_mva))
        for d in data.get('lines', []):
            self.add_line(Line(param_dict=d, s_n_sys=self.base_mva, v_n_sys=self.base_voltage))
        for d in data.get('transformers', []):
            self.add_transformer(Transformer(param_dict=d, s_n_sys=self.base_mva))
        for key, sub, cls, adder in (('avr', 'SEXS', SEXS, self.add_exciter), ('gov', 'TGOV1', TGOV1, self.add_governor),
                                     ('pss', 'STAB1', STAB1, self.add_pss)):
            table = data.get(key, [])
            if isinstance(table, dict):
                for d in table[sub]:
                    adder(cls(param_dict=d))
        self.set_slack_bus(data['slack_bus'])

    def set_slack_bus(self, slack_bus):
        self.busses[self.bus_idxs[slack_bus]].lf_type = 'SL'

    def _par(self, model):
        model.enable_parallel_simulation(self.parallel_sims, self.prologue_device)
        return model

    def add_bus(self, bus_model):
        self.bus_idxs[bus_model.name] = len(self.busses)
        self.busses.append(self._par(bus_model))

    def _bus_of(self, model):
        return self.busses[self.bus_idxs[model.bus]]

    def add_generator(self, generator_model):
        bus = self._bus_of(generator_model)
        bus.add_model(self._par(generator_model))
        bus.voltage = generator_model.v_soll          # simulator.py:244
        bus.lf_type = 'PV'

    def add_load(self, load_model):
        self._bus_of(load_model).add_model(self._par(load_model))

    def add_shunt(self, shunt_model):
        self._bus_of(shunt_model).add_model(self._par(shunt_model))

    def _add_branch(self, model, store):
        model.from_bus_id = self.bus_idxs[model.from_bus_name]
        model.to_bus_id = self.bus_idxs[model.to_bus_name]
        store.append(self._par(model))

    def add_line(self, line_model):
        self._add_branch(line_model, self.lines)

    def add_transformer(self, transformer_model):
        self._add_branch(transformer_model, self.trafos)

    def add_exciter(self, exciter_model):
        generator = self.get_generator_by_name(exciter_model.gen)
        exciter_model.v_setpoint = generator.v_soll   # simulator.py:319
        generator.add_exciter(self._par(exciter_model))

    def add_governor(self, governor_model):
        self.get_generator_by_name(governor_model.gen).add_governor(self._par(governor_model))

    def add_pss(self, pss_model):
        self.get_generator_by_name(pss_model.gen).add_pss(self._par(pss_model))

    def add_inverter(self, inverter_model):
        """simulator.py:247-252: the model is attached to its bus like any other.  The reference ships no inverter model
        (only the PIController block such a model would use, blocks.py:7-100), so there are no equations the stepping
        kernels could integrate: ``run()`` refuses a bus model it does not know instead of silently ignoring it."""
        bus = self._bus_of(inverter_model)
        try:
            inverter_model.enable_parallel_simulation(self.parallel_sims, self.prologue_device)
        except TypeError:
            inverter_model.enable_parallel_simulation(self.parallel_sims)       # reference signature
        bus.add_model(inverter_model)

    def add_sc_event(self, start_time, end_time, bus):
        self.sc_events.append(ScEvent(start_time, end_time, self.bus_idxs[bus]))

    def add_param_event(self, timestep, model, parameter, new_val):
        self.param_events.append(ParamEvent(timestep, model, parameter, new_val))

    def set_record_function(self, record_func):
        self.record_func = record_func
        self.record_channels = None

    def set_record_channels(self, names):
        """Extension: recorder channels by name ('G1.omega', 'bus:B8.v_re', ...) instead of a Python callable."""
        self.record_channels = list(names)
        self.record_func = None

    def invalidate_caches(self):
        """Drop every memoised prologue result (load flow, inv(Y_dyn), derived model constants, packed tables, event
        schedule).  The memoisation keys tensors on identity + version counter; a write through ``.data``
        (``p.data += ...``, as the reference's own optimiser does, optimization_lib/optimizers.py:83) bumps neither, so
        after such a write on a tensor that does NOT require grad call this before the next run().  (Tensors that require
        grad always bypass the caches.)"""
        for k in ('_prologue_cache', '_schedule_cache'):
            self.__dict__.pop(k, None)
        self.__dict__.setdefault('_pack_cache', {}).clear()
        self._uniform_cache = {}

    def reset(self):
        """simulator.py:485-495 / static_models.py:189-197: bus voltages back to the last model's set point (or 1.0),
        cached matrices dropped.  The next run() redoes load flow and initialisation from scratch."""
        P = self.parallel_sims
        if getattr(self, '_flat_start', None) is None:
            # one shared (never written) tensor, so that the prologue's memoisation sees the same input after every reset
            self._flat_start = mark_uniform(torch.ones((P, 1), dtype=torch.complex128, device=self.prologue_device))
        for bus in self.busses:
            v = getattr(bus.models[-1], 'v_soll', None) if bus.models else None
            bus.voltage = v if v is not None else self._flat_start
        self.static_y_matrix = None
        self.inverse_dynamic_y_matrix = None

    # ------------------------------------------------------------------------------------------------------------
    # prologue + time loop
    # ------------------------------------------------------------------------------------------------------------
    def initialize(self):
        """Load flow, model initialisation, Z = inv(Y_dyn), initial bus voltages (simulator.py:438-450)."""
        for bus in self.busses:
            for m in bus.models:
                if not isinstance(m, (SynchMachine, Load, Shunt)):
                    raise NotImplementedError("bus model %r on %s has no kernel implementation: backend='cuda' integrates "
                                              'SynchMachine (+ SEXS / TGOV1 / STAB1), constant-impedance loads and shunts'
                                              % (type(m).__name__, bus.name))
        # kernels of libdps_b200.so when nothing needs a gradient through the prologue (prologue_device.py), else the
        # differentiable torch restatement
        init = None
        if prologue_device.applicable(self, self.prologue_device):
            init = prologue_device.initialize(self, self.prologue_device)
        self.last_prologue = 'kernels' if init is not None else 'torch'
        if init is None:
            init = prologue.initialize(self, self.prologue_device)
        self.inverse_dynamic_y_matrix = init.Z0
        P = self.parallel_sims
        packed = getattr(init, 'packed', None)
        if packed is not None:
            inj = packed['inj_bus']                                               # [B,n] from the kernel path
        else:
            cur = prologue.initial_injections(init)                               # [B,G]
            B = max(cur.shape[0], init.Z0.shape[0])
            n = len(self.busses)
            inj = torch.zeros((B, n), dtype=torch.complex128, device=self.prologue_device)
            cols = [inj[:, i] for i in range(n)]
            for gi, b in enumerate(init.gen_bus):
                cols[b] = cols[b] + cur[:, gi]
            inj = torch.stack(cols, dim=1)
        volt = torch.matmul(init.Z0, inj.unsqueeze(-1)).squeeze(-1)               # simulator.py:447
        for i, bus in enumerate(self.busses):
            bus.voltage = volt[:, i:i + 1].expand(P, 1)
        self.last_init = init
        return init

    def _checkpoint_stride(self, n_gen):
        """Checkpoint stride of a run that needs gradients.  ``ckpt_stride`` given by the caller wins.  Otherwise 1 -- the
        state before every step is kept and the adjoint evaluates each step's stages in registers, without re-integration
        or scratch -- for systems of up to 4 generators whose n_steps x n_gen x 13 x P doubles fit a quarter of the free
        device memory (at most 24 GB); else the shortest of 8, 16, 32 whose checkpoints fit that budget.  With more generators a warp holds few trajectories, its checkpoint stores
        are 24..48-byte pieces and writing every step costs the forward pass more than the adjoint gains (IEEE-39, 8192
        sims: forward +11 ms).  A short stride keeps the per-segment scratch of the adjoint (every stage's inputs, sin/cos
        and bus voltage for each step of a segment) near L2 size: IEEE-39, 8192 sims: 25 ms per backward pass at strides 4 and
        8, ~27 ms at 16; the forward pass does not notice (profiles/round2/adjoint_r3.txt)."""
        if self.ckpt_stride is not None:
            return int(self.ckpt_stride)
        full = len(self.time) * n_gen * engine.N_STATES * self.parallel_sims * 8
        budget = 24e9
        if torch.cuda.is_available() and str(self.device).startswith('cuda'):
            free, _total = torch.cuda.mem_get_info(torch.device(self.device))
            budget = min(budget, free / 4)
        if n_gen <= 4 and full <= budget:
            return 1
        for stride in (8, 16):
            if full / stride <= budget:
                return stride
        return 32

    def _trace_recorder(self, init):
        if self.record_channels is not None:
            return recording.from_names(self, init.gens, self.record_channels)
        if self.record_func is None:
            raise TypeError('No record function specified')
        return recording.trace(self, init.gens, self.record_func)

    def prepare(self):
        """Everything ``run()`` does before the time loop: prologue, recorder tracing, planning, packing.
        Returns a :class:`PreparedRun` whose ``launch()`` enqueues the kernels (used by bench.py to time the loop
        alone, as the reference's own timer does -- simulator.py:507, 548-550)."""
        init = self.initialize()
        P, dev, pdev = self.parallel_sims, self.device, self.prologue_device
        rplan = self._trace_recorder(init)
        packed = getattr(init, 'packed', None)             # kernel prologue: inputs already in the stepping kernel's layout
        net_grad = torch.is_grad_enabled() and (init.Z0.requires_grad or init.Ydiag.requires_grad)
        needs_grad = packed is None and (net_grad or (torch.is_grad_enabled() and (
            any(t.requires_grad for x in init.x0 for t in x) or
            any(v.requires_grad for sl in init.slots for v in sl.values()))))
        topo, extras = planning.build_topology(self, init, rplan.channel_keys, self.solver,
                                               self._checkpoint_stride(len(init.gens)) if needs_grad else 0)
        pack_cache = self.__dict__.setdefault('_pack_cache', {})
        if packed is not None:
            G = len(init.gens)
            x0 = packed['x0'] if packed['x0'].shape[-1] == P else packed['x0'].expand(G, engine.N_STATES, P).contiguous()
            upar = packed['slots'][:, :, 0].contiguous()
            vpar = None if packed['vrows'] is None else packed['slots'].reshape(G * engine.N_SLOTS, -1).index_select(0, packed['vrows'])
            slot_vary = packed['slot_vary']
            topo['common_ctrl'] = planning.common_controller_constants_keyed(init.ctrl, upar, slot_vary, packed['ctrl_key'],
                                                                              pack_cache)
        else:
            x0 = planning.canonical_x0(init, P)
            upar, vpar, slot_vary = planning.pack_slots(init, P, pdev, pack_cache)
            topo['common_ctrl'] = planning.common_controller_constants(init, upar, slot_vary, pack_cache)
        topo['slot_vary'] = slot_vary
        topo['n_vary'] = 0 if vpar is None else vpar.shape[0]
        topo['want'] = adjoint.wanted_slots(init, prologue.SLOTS) if needs_grad else None
        topo['want_z'] = bool(net_grad)
        plan = engine.Plan(topo)
        self.last_plan = plan
        with torch.no_grad():
            z_re, z_im = planning.split_complex_batch_last(init.Z0)
            yd_re, yd_im = planning.split_complex_batch_last(init.Ydiag)
            if pdev != dev:                              # prologue evaluated elsewhere: upload the packed kernel inputs
                z_re, z_im, yd_re, yd_im = (t.to(dev) for t in (z_re, z_im, yd_re, yd_im))
            zsub = torch.empty((topo['n_cfg'], topo['z_rows'], topo['z_cols'], 2, z_re.shape[-1]),
                               dtype=torch.float64, device=dev)
        if pdev != dev:                                  # .to() is differentiable: gradients flow back to the leaves
            x0, upar = x0.to(dev), upar.to(dev)
            vpar = None if vpar is None else vpar.to(dev)
        return PreparedRun(self, init, rplan, plan, topo, extras, x0, upar, vpar, (z_re, z_im, yd_re, yd_im), zsub)

    def run(self, host_out=None, stream_chunks=8):
        """Initialise, integrate every step of every trajectory on the GPU, return (time, result[P,T,n]).

        ``host_out`` (extension): a pinned host tensor ``[n_values, T, P]``; the recorder is then streamed to it in
        ``stream_chunks`` time slices while later steps are still being integrated, and the returned result is a view of
        ``host_out`` (no autograd)."""
        t0 = _time.time()
        prep = self.prepare()
        t1 = _time.time()
        pending = [ev for ev in self.param_events if not ev.handled]
        if pending or host_out is not None:
            result, x_final = prep.execute_ranges(pending, host_out, stream_chunks)
        else:
            result, x_final = prep.execute()
        if self.verbose:
            torch.cuda.synchronize()
            print('Dynamic simulation finished in {:.2f} seconds'.format(_time.time() - t1))
        self.timings = {'prologue_s': t1 - t0}
        self._publish_final_state(prep.init, x_final)
        return self.time, result

    def _publish_final_state(self, init, x_final):
        """Leave the post-run states on the model objects, as the reference's set_state_vector does."""
        P = self.parallel_sims
        xf = x_final.detach()
        names = SynchMachine.STATES
        for gi, g in enumerate(init.gens):
            for k, n in enumerate(names):
                setattr(g, n, torch.complex(xf[gi, k], torch.zeros_like(xf[gi, k])).reshape(P, 1))
            if g.exciter:
                g.exciter.lead_lag.state_1 = xf[gi, 6].reshape(P, 1)
                g.exciter.pt1_lim.state_1 = xf[gi, 7].reshape(P, 1)
            if g.governor:
                g.governor.pt1_lim.state_1 = xf[gi, 8].reshape(P, 1)
                g.governor.lead_lag.state_1 = xf[gi, 9].reshape(P, 1)
            if g.stabilizer:
                g.stabilizer.washout.state_1 = xf[gi, 10].reshape(P, 1)
                g.stabilizer.lead_lag1.state_1 = xf[gi, 11].reshape(P, 1)
                g.stabilizer.lead_lag2.state_1 = xf[gi, 12].reshape(P, 1)


class PreparedRun(object):
    """Device-resident inputs of one run plus the two kernel launches of the time loop."""

    def __init__(self, sim, init, rplan, plan, topo, extras, x0, upar, vpar, zparts, zsub):
        self.sim, self.init, self.rplan, self.plan, self.topo, self.extras = sim, init, rplan, plan, topo, extras
        self.x0, self.upar, self.vpar, self.zparts, self.zsub = x0, upar, vpar, zparts, zsub
        self.kernels_per_launch = 2          # dps_network_reduce_kernel + dps_step_kernel

    def reduce_network(self):
        """On-device fault re-reduction (replaces the per-step double inversion of simulator.py:526-528)."""
        z_re, z_im, yd_re, yd_im = self.zparts
        engine.network_reduce(z_re, z_im, yd_re, yd_im, self.extras['obs_bus'], self.extras['gen_bus'],
                              self.extras['cfg_sets'], planning.FAULT_ADMITTANCE, self.zsub)

    def launch(self, rec, x_final=None):
        """Raw, autograd-free time loop into caller-provided buffers (both kernels, current stream)."""
        self.reduce_network()
        self.plan.forward(self.sim.parallel_sims, self.x0.detach(), self.upar.detach(),
                          None if self.vpar is None else self.vpar.detach(), self.zsub, rec, x_final, None)

    def execute(self):
        engine.require_cuda()
        if self.topo.get('want_z'):
            # a gradient flows into the network: the fault re-reduction becomes part of the torch graph
            zsub = planning.network_reduce_torch(self.init.Z0, self.init.Ydiag, self.extras['obs_bus'],
                                                 self.extras['gen_bus'], self.extras['cfg_sets'], self.extras['m_max'])
            zsub = zsub.to(self.sim.device)
        else:
            with torch.no_grad():
                self.reduce_network()
            zsub = self.zsub
        rec, x_final = StepFunction.apply(self.plan, self.sim.parallel_sims, self.x0, self.upar, self.vpar, zsub)
        return self._assemble(rec), x_final.detach()

    def _needs_grad(self):
        topo = self.topo
        return torch.is_grad_enabled() and (topo.get('ckpt_stride', 0) > 0 or topo.get('want_z') or self.x0.requires_grad or
                                            self.upar.requires_grad or (self.vpar is not None and self.vpar.requires_grad))

    def execute_ranges_grad(self, param_events):
        """Parameter events WITH gradients: every step range between two events is its own plan (its slice of the fault
        schedule, its constants, its checkpoints) and its own node of the autograd graph; the final state of a range is the
        differentiable initial state of the next, so the co-state travels back through the events exactly as torch
        autograd carries it through the reference's loop (a parameter that an event overwrites only collects the gradient
        of the steps before the event)."""
        engine.require_cuda()
        sim, topo = self.sim, self.topo
        P, T, dev = sim.parallel_sims, topo['n_steps'], sim.device
        cuts, when = {0, T}, {}
        for ev in param_events:
            hit = [k for k, t in enumerate(sim.time) if t >= ev.start_time]      # ParamEvent.handle_event: t >= start
            if hit:
                cuts.add(hit[0])
                when.setdefault(hit[0], []).append(ev)
        cuts = sorted(cuts)
        if topo.get('want_z'):
            zsub = planning.network_reduce_torch(self.init.Z0, self.init.Ydiag, self.extras['obs_bus'],
                                                 self.extras['gen_bus'], self.extras['cfg_sets'], self.extras['m_max']).to(dev)
        else:
            with torch.no_grad():
                self.reduce_network()
            zsub = self.zsub
        stride = topo.get('ckpt_stride', 0) or sim._checkpoint_stride(topo['n_gen'])
        x_cur, upar, vpar = self.x0, self.upar, self.vpar
        slot_vary, want, common = topo['slot_vary'], topo.get('want'), topo.get('common_ctrl')
        if want is None:
            want = adjoint.wanted_slots(self.init, prologue.SLOTS)
        recs = []
        for k0, k1 in zip(cuts[:-1], cuts[1:]):
            if k0 in when:
                for ev in when[k0]:
                    setattr(ev.model, ev.param_name, ev.value)
                    ev.handled = True
                saved = self.init.slots
                self.init.slots = prologue.refresh_loop_slots(sim, self.init, sim.prologue_device)
                try:
                    upar, vpar, slot_vary = planning.pack_slots(self.init, P, sim.prologue_device)
                    want = adjoint.wanted_slots(self.init, prologue.SLOTS)
                    common = planning.common_controller_constants(self.init, upar, slot_vary)
                finally:
                    self.init.slots = saved
                upar, vpar = upar.to(dev), None if vpar is None else vpar.to(dev)
            t_r = dict(topo, n_steps=k1 - k0, step_cfg=list(topo['step_cfg'][k0:k1]), slot_vary=slot_vary,
                       n_vary=0 if vpar is None else vpar.shape[0], ckpt_stride=stride, want=want, common_ctrl=common)
            rec_r, x_cur = StepFunction.apply(engine.Plan(t_r), P, x_cur, upar, vpar, zsub)
            recs.append(rec_r)
        return self._assemble(torch.cat(recs, dim=1)), x_cur.detach()

    def _assemble(self, rec):
        rows = self.extras['key_rows']
        if rows != list(range(rec.shape[0])):
            rec = rec[torch.tensor(rows, device=rec.device)]
        return self.rplan.assemble(rec)

    def execute_ranges(self, param_events, host_out=None, stream_chunks=8):
        """The time loop as a sequence of step ranges (``dps_forward_range``), for two purposes:

        * **parameter events** (``ParamEvent``, static_models.py:52-95; applied at simulator.py:533-534 before the solver
          step of the first grid time ``t >= start_time``): the run is cut at those steps, the attribute is set, the
          constant slots are re-derived and the next range continues from the carried state.  Like the reference, nothing
          is re-initialised and Z is not rebuilt;
        * **streaming the recorder to the host**: each finished range is copied to ``host_out`` on a second stream while
          the next range is being integrated.
        Parameter events with gradients go through :meth:`execute_ranges_grad`; a streamed run carries no autograd."""
        engine.require_cuda()
        sim, topo = self.sim, self.topo
        if self._needs_grad():
            # ckpt_stride > 0 <=> prepare() saw a tensor requiring grad (a model parameter, or a network parameter that
            # reaches Z0 / Ydiag)
            if host_out is None:
                return self.execute_ranges_grad(param_events)
            # the streamed ranges carry no autograd graph: refuse instead of returning a detached result
            raise NotImplementedError('gradients through a run streamed to the host are not supported: wrap run() in '
                                      'torch.no_grad() or drop requires_grad')
        P, T, dev = sim.parallel_sims, topo['n_steps'], sim.device
        cuts = {0, T}
        when = {}
        for ev in param_events:
            hit = [k for k, t in enumerate(sim.time) if t >= ev.start_time]      # ParamEvent.handle_event: t >= start
            if hit:
                cuts.add(hit[0])
                when.setdefault(hit[0], []).append(ev)
        if host_out is not None:
            n = max(int(stream_chunks), 1)
            cuts |= {(T * i) // n for i in range(1, n)}
        cuts = sorted(cuts)
        with torch.no_grad():
            self.reduce_network()
            rec = torch.empty((len(topo['rec']), T, P), dtype=torch.float64, device=dev)
            x_cur = self.x0.detach().clone()
            plan, upar, vpar = self.plan, self.upar.detach(), None if self.vpar is None else self.vpar.detach()
            slots = self.init.slots
            copy_stream = torch.cuda.Stream() if host_out is not None else None
            main = torch.cuda.current_stream()
            rows = self.extras['key_rows']
            for k0, k1 in zip(cuts[:-1], cuts[1:]):
                if k0 in when:
                    for ev in when[k0]:
                        setattr(ev.model, ev.param_name, ev.value)
                        ev.handled = True
                    slots = prologue.refresh_loop_slots(sim, self.init, sim.prologue_device)
                    saved, self.init.slots = self.init.slots, slots
                    upar, vpar, slot_vary = planning.pack_slots(self.init, P, sim.prologue_device)
                    upar, vpar = upar.to(dev), None if vpar is None else vpar.to(dev)
                    t2 = dict(topo, slot_vary=slot_vary, n_vary=0 if vpar is None else vpar.shape[0], ckpt_stride=0,
                              want=None)
                    t2['common_ctrl'] = planning.common_controller_constants(self.init, upar, slot_vary)
                    self.init.slots = saved
                    plan = engine.Plan(t2)
                plan.forward_range(P, k0, k1, x_cur, upar, vpar, self.zsub, rec, x_cur, None)
                if host_out is not None:
                    done = torch.cuda.Event()
                    done.record(main)
                    copy_stream.wait_event(done)
                    with torch.cuda.stream(copy_stream):
                        if (rows == list(range(rec.shape[0])) and host_out.is_contiguous() and host_out.is_pinned()
                                and not os.environ.get('DPS_ROW_COPIES')):
                            # one strided copy per time slice (all channels) instead of one per channel
                            engine._check(plan.lib, plan.lib.dps_recorder_to_host(
                                ctypes.c_void_p(host_out.data_ptr()), engine._ptr(rec), rec.shape[0], T * P, k0 * P, (k1 - k0) * P,
                                ctypes.c_void_p(copy_stream.cuda_stream)), 'dps_recorder_to_host')
                        else:
                            for i, r in enumerate(rows):
                                host_out[i, k0:k1].copy_(rec[r, k0:k1], non_blocking=True)
            if host_out is not None:
                main.wait_stream(copy_stream)
                torch.cuda.synchronize()
                index = {k: i for i, k in enumerate(self.rplan.channel_keys)}
                if not self.rplan.direct:
                    return self.rplan.assemble(host_out), x_cur
                return host_out.permute(2, 1, 0), x_cur
        return self._assemble(rec), x_cur


Pss = PowerSystemSimulation
