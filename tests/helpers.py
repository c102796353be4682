"""Shared test helpers: building simulations from the golden fixtures, the host emulation runner, oracle inputs."""
import ctypes
import json
import os
import subprocess

import numpy as np
import torch

from diffpssi_b200 import Pss, engine, planning
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')
EMUL_SRC = os.path.join(ROOT, 'tests', 'host_emul', 'emul.cpp')
EMUL_LIB = os.path.join(ROOT, 'tests', 'host_emul', 'libdps_emul.so')
DT = 0.005


def load_grid(name):
    with open(os.path.join(GOLD, 'grids', name + '.json')) as f:
        return json.load(f)


def load_gold(case):
    return np.load(os.path.join(GOLD, case + '.npz'))


def resolve_attr(sim, gen_name, attr):
    obj = sim.get_generator_by_name(gen_name)
    parts = attr.split('.')
    for p in parts[:-1]:
        obj = getattr(obj, p)
    return obj, parts[-1]


def apply_overrides(sim, overrides, requires_grad=False, complex_leaves=True):
    """overrides: {(gen name, attr path): array[P]} -> sets [P,1] tensors like the reference examples do
    (examples/models/ibb_model/ibb_opt.py:48-57).  Returns the leaf tensors."""
    leaves = {}
    for (gname, attr), val in overrides.items():
        obj, name = resolve_attr(sim, gname, attr)
        arr = np.asarray(val, dtype=np.float64).reshape(-1, 1)
        is_limit = name in ('limit', 'lim_min', 'lim_max')
        old = getattr(obj, name)
        want_c = (torch.is_tensor(old) and old.is_complex()) or (requires_grad and complex_leaves and not is_limit)
        t = torch.tensor(arr, dtype=torch.complex128 if want_c else torch.float64, device=sim.device,
                         requires_grad=requires_grad and not is_limit)
        setattr(obj, name, t)
        if not is_limit:
            leaves[(gname, attr)] = t
    return leaves


def gold_overrides(gold):
    out = {}
    for key in gold.files:
        if key.startswith('override/'):
            _k, gname, attr = key[len('override/'):].split('|')
            out[(gname, attr)] = gold[key]
    return out


def sim_from_gold(case, grid_name, device=None, requires_grad=False, sim_time=None, P=None, tile=1):
    """Simulation configured like the golden case (grid, solver, events, per-sim overrides, all golden channels)."""
    gold = load_gold(case)
    grid = load_grid(grid_name)
    P0 = gold['traj'].shape[0]
    P = (P or P0) * tile
    t = gold['time']
    sim = Pss(parallel_sims=P, sim_time=sim_time if sim_time is not None else len(t) * DT, time_step=DT,
              solver=str(gold['solver']), grid_data=grid, verbose=False, device=device)
    assert len(sim.time) == len(t) or sim_time is not None
    for (t0, t1, b) in json.loads(str(gold['events'])):
        sim.add_sc_event(t0, t1, b)
    ov = {k: np.tile(v, tile) for k, v in gold_overrides(gold).items()}
    leaves = apply_overrides(sim, ov, requires_grad=requires_grad)
    if 'param_events' in gold.files:
        for t_ev, path, value in json.loads(str(gold['param_events'])):
            gname, attr = path.split('|', 1)
            obj, name = resolve_attr(sim, gname, attr)
            old = getattr(obj, name)
            dt_ = old.dtype if torch.is_tensor(old) else torch.float64
            sim.add_param_event(t_ev, obj, name, torch.ones((P, 1), dtype=dt_, device=sim.device) * value)
    sim.set_record_channels([str(s) for s in gold['names']])
    return sim, gold, leaves


# ----------------------------------------------------------------------------------------------------------------------
# host emulation of the device code (tests/host_emul/emul.cpp)
# ----------------------------------------------------------------------------------------------------------------------
def build_emul():
    deps = [EMUL_SRC] + [os.path.join(ROOT, 'diffpssi_b200', 'csrc', f) for f in
                         ('dps_model.cuh', 'dps_reduce.cuh', 'dps_internal.h')] + [os.path.join(ROOT, 'include', 'dps.h')]
    if not os.path.exists(EMUL_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMUL_LIB) for d in deps):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-shared', '-fPIC', '-o', EMUL_LIB, EMUL_SRC])
    return ctypes.CDLL(EMUL_LIB)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def run_emulated(sim):
    """sim.run() with the two kernels replaced by their host emulation.  Exercises the real prologue, recorder tracing
    and planning code of the package.  Returns (result [P,T,n] ndarray, dict of intermediates)."""
    lib = build_emul()
    init = sim.initialize()
    P = sim.parallel_sims
    rplan = sim._trace_recorder(init)
    topo, extras = planning.build_topology(sim, init, rplan.channel_keys, sim.solver, 0)
    x0 = planning.canonical_x0(init, P).detach().cpu().numpy()
    upar_t, vpar_t, slot_vary = planning.pack_slots(init, P, sim.device)
    upar = upar_t.detach().cpu().numpy()
    vpar = vpar_t.detach().cpu().numpy() if vpar_t is not None else None
    z_re, z_im = (a.detach().cpu().numpy() for a in planning.split_complex_batch_last(init.Z0))
    yd_re, yd_im = (a.detach().cpu().numpy() for a in planning.split_complex_batch_last(init.Ydiag))
    Bz = z_re.shape[-1]
    zsub = np.zeros((topo['n_cfg'], topo['z_rows'], topo['z_cols'], 2, Bz))
    ci = ctypes.c_int
    ptr, flat = [0], []
    for s in extras['cfg_sets']:
        flat += list(s)
        ptr.append(len(flat))
    lib.emul_network_reduce(ctypes.c_longlong(Bz), len(sim.busses), _p(z_re), _p(z_im), _p(yd_re), _p(yd_im),
                            len(extras['obs_bus']), (ci * len(extras['obs_bus']))(*extras['obs_bus']),
                            len(extras['gen_bus']), (ci * len(extras['gen_bus']))(*extras['gen_bus']),
                            len(extras['cfg_sets']), (ci * len(ptr))(*ptr), (ci * max(len(flat), 1))(*flat),
                            ctypes.c_double(planning.FAULT_ADMITTANCE), _p(zsub))
    T, G = topo['n_steps'], topo['n_gen']
    chans = topo['rec']
    rec = np.zeros((len(chans), T, P))
    x_final = np.zeros((G, engine.N_STATES, P))
    step_cfg = np.asarray(topo['step_cfg'], dtype=np.uint8)
    flat_ch = [v for c in chans for v in c]
    # parameter events cut the run into ranges exactly like PreparedRun.execute_ranges does on the GPU
    from diffpssi_b200 import prologue as _prologue
    cuts, when = {0, T}, {}
    for ev in [e for e in sim.param_events if not e.handled]:
        hit = [k for k, t in enumerate(sim.time) if t >= ev.start_time]
        if hit:
            cuts.add(hit[0])
            when.setdefault(hit[0], []).append(ev)
    cuts = sorted(cuts)
    x_cur = x0
    for k0, k1 in zip(cuts[:-1], cuts[1:]):
        if k0 in when:
            for ev in when[k0]:
                setattr(ev.model, ev.param_name, ev.value)
                ev.handled = True
            saved, init.slots = init.slots, _prologue.refresh_loop_slots(sim, init, sim.device)
            upar_t, vpar_t, slot_vary = planning.pack_slots(init, P, sim.device)
            init.slots = saved
            upar = upar_t.detach().cpu().numpy()
            vpar = vpar_t.detach().cpu().numpy() if vpar_t is not None else None
        n = k1 - k0
        rec_part = np.zeros((len(chans), n, P))
        cfg_part = np.ascontiguousarray(step_cfg[k0:k1])
        lib.emul_forward(G, topo['solver'], n, ctypes.c_double(topo['dt']), ctypes.c_double(topo['fn']),
                         (ci * G)(*topo['ctrl']), topo['n_obs'], topo['z_rows'], topo['z_cols'], topo['n_cfg'],
                         _p(cfg_part), topo['z_per_sim'], (ci * len(slot_vary))(*slot_vary), len(chans),
                         (ci * max(len(flat_ch), 1))(*flat_ch), ctypes.c_longlong(P), _p(np.ascontiguousarray(x_cur)),
                         _p(upar), _p(vpar), _p(zsub), _p(rec_part), _p(x_final))
        rec[:, k0:k1] = rec_part
        x_cur = x_final.copy()
    result = rplan.assemble(torch.from_numpy(rec[extras['key_rows']])).numpy()
    return result, dict(init=init, topo=topo, extras=extras, zsub=zsub, x0=x0, upar=upar, vpar=vpar, rec=rec,
                        x_final=x_final, slot_vary=slot_vary)


from oracle.adapter import oracle_inputs_from_sim, run_oracle_for  # noqa: E402,F401


# ----------------------------------------------------------------------------------------------------------------------
# emulated autograd path (CPU): the package's prepare() + host builds of the forward / adjoint device code
# ----------------------------------------------------------------------------------------------------------------------
class EmulStepFunction(torch.autograd.Function):
    """CPU stand-in for diffpssi_b200.autograd.StepFunction used by the `not gpu` gradient tests."""

    @staticmethod
    def forward(ctx, prep, x0, upar, vpar, zsub):
        lib = build_emul()
        zsub_np = np.ascontiguousarray(zsub.detach().numpy()) if torch.is_tensor(zsub) else zsub
        topo = prep.topo
        P = prep.sim.parallel_sims
        T, G = topo['n_steps'], topo['n_gen']
        ci = ctypes.c_int
        chans = topo['rec']
        flat_ch = [v for c in chans for v in c]
        args = dict(
            head=(G, topo['solver'], T, ctypes.c_double(topo['dt']), ctypes.c_double(topo['fn']),
                  (ci * G)(*topo['ctrl']), topo['n_obs'], topo['z_rows'], topo['z_cols'], topo['n_cfg'],
                  _p(np.asarray(topo['step_cfg'], dtype=np.uint8)), topo['z_per_sim'],
                  (ci * len(topo['slot_vary']))(*topo['slot_vary']), len(chans), (ci * max(len(flat_ch), 1))(*flat_ch),
                  ctypes.c_longlong(P)))
        x0n = np.ascontiguousarray(x0.detach().numpy())
        upn = np.ascontiguousarray(upar.detach().numpy())
        vpn = np.ascontiguousarray(vpar.detach().numpy()) if vpar is not None else None
        rec = np.zeros((len(chans), T, P))
        xf = np.zeros((G, engine.N_STATES, P))
        keep = np.asarray(topo['step_cfg'], dtype=np.uint8)
        head = list(args['head'])
        head[10] = _p(keep)
        lib.emul_forward(*head, _p(x0n), _p(upn), _p(vpn), _p(zsub_np), _p(rec), _p(xf))
        ctx.saved = (head, keep, x0n, upn, vpn, zsub_np, topo, P)
        return torch.from_numpy(rec)

    @staticmethod
    def backward(ctx, g_rec):
        lib = build_emul()
        head, keep, x0n, upn, vpn, zsub_np, topo, P = ctx.saved
        G = topo['n_gen']
        g = np.ascontiguousarray(g_rec.numpy())
        g_x0 = np.zeros((G, engine.N_STATES, P))
        gslot = np.zeros((G, engine.N_SLOTS, P))
        want_z = ctx.needs_input_grad[4]
        P_ = g_x0.shape[-1]
        gz = np.zeros(zsub_np.shape[:4] + (P_,)) if want_z else None
        lib.emul_backward(*head, _p(x0n), _p(upn), _p(vpn), _p(zsub_np), _p(g), _p(g_x0), _p(gslot), _p(gz))
        g_upar = torch.zeros(upn.shape, dtype=torch.float64)
        g_vpar = torch.zeros(vpn.shape, dtype=torch.float64) if vpn is not None else None
        for i, vi in enumerate(topo['slot_vary']):
            gi, k = divmod(i, engine.N_SLOTS)
            if vi >= 0:
                g_vpar[vi] = torch.from_numpy(gslot[gi, k])
            else:
                g_upar[gi, k] = float(gslot[gi, k].sum())
        g_z = None
        if want_z:
            g_z = torch.from_numpy(gz if zsub_np.shape[-1] > 1 else gz.sum(axis=-1, keepdims=True))
        return None, torch.from_numpy(g_x0), g_upar, g_vpar, g_z


def run_emulated_autograd(sim):
    """Differentiable CPU run: the package's real prologue graph + emulated kernels.  Returns result [P,T,n] tensor."""
    prep = sim.prepare()
    lib = build_emul()
    topo, extras = prep.topo, prep.extras
    z_re, z_im, yd_re, yd_im = (a.detach().cpu().numpy() for a in prep.zparts)
    Bz = z_re.shape[-1]
    zsub = np.zeros((topo['n_cfg'], topo['z_rows'], topo['z_cols'], 2, Bz))
    ci = ctypes.c_int
    ptr, flat = [0], []
    for s in extras['cfg_sets']:
        flat += list(s)
        ptr.append(len(flat))
    lib.emul_network_reduce(ctypes.c_longlong(Bz), len(sim.busses), _p(z_re), _p(z_im), _p(yd_re), _p(yd_im),
                            len(extras['obs_bus']), (ci * len(extras['obs_bus']))(*extras['obs_bus']),
                            len(extras['gen_bus']), (ci * len(extras['gen_bus']))(*extras['gen_bus']),
                            len(extras['cfg_sets']), (ci * len(ptr))(*ptr), (ci * max(len(flat), 1))(*flat),
                            ctypes.c_double(planning.FAULT_ADMITTANCE), _p(zsub))
    if topo.get('want_z'):
        zsub = planning.network_reduce_torch(prep.init.Z0, prep.init.Ydiag, extras['obs_bus'], extras['gen_bus'],
                                             extras['cfg_sets'], extras['m_max'])
    rec = EmulStepFunction.apply(prep, prep.x0, prep.upar, prep.vpar, zsub)
    rows = extras['key_rows']
    if rows != list(range(rec.shape[0])):
        rec = rec[torch.tensor(rows)]
    return prep.rplan.assemble(rec)
