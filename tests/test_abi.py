"""The C-ABI library loads and exports every symbol include/dps.h declares; argument validation works without a GPU."""
import ctypes
import os
import re

import pytest

from diffpssi_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, 'include', 'dps.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dps_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    names = declared_functions()
    assert set(names) == set(engine.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.dps_abi_version() == engine.ABI_VERSION


def test_plan_validation_errors_are_reported():
    lib = engine.load_library()
    topo = dict(n_gen=2, solver=1, n_steps=4, dt=0.005, fn=60.0, ctrl=[0, 0], n_obs=2, z_rows=2, z_cols=2, n_cfg=1,
                step_cfg=[0, 0, 0, 0], z_per_sim=0, slot_vary=[-1] * (2 * engine.N_SLOTS), n_vary=0,
                rec=[(0, 1, 0), (0, 0, 13)])
    plan = engine.Plan(topo)
    assert plan.n_checkpoints() == 0
    bad = dict(topo, rec=[(0, 5, 0)])
    with pytest.raises(engine.EngineError, match='out of range'):
        engine.Plan(bad)
    with pytest.raises(engine.EngineError, match='duplicates'):
        engine.Plan(dict(topo, rec=[(0, 1, 0), (0, 1, 0)]))
    # 17..128 generators select the wide stepper; more are rejected, and so are the things the wide stepper cannot do
    wide = dict(topo, n_gen=40, ctrl=[0] * 40, slot_vary=[-1] * (40 * engine.N_SLOTS), n_obs=40, z_rows=40, z_cols=40)
    assert engine.Plan(wide).handle
    with pytest.raises(engine.EngineError, match='stepper limit'):
        engine.Plan(dict(topo, n_gen=200, ctrl=[0] * 200, slot_vary=[-1] * (200 * engine.N_SLOTS), n_obs=200, z_rows=200,
                         z_cols=200))
    with pytest.raises(engine.EngineError, match='shared by the batch'):
        engine.Plan(dict(wide, z_per_sim=1))
    assert engine.Plan(dict(wide, ckpt_stride=8)).n_checkpoints() == 1        # the wide stepper has an adjoint since round 2
    with pytest.raises(engine.EngineError, match='n_cfg'):
        engine.Plan(dict(topo, step_cfg=[0, 3, 0, 0]))
    ck = engine.Plan(dict(topo, ckpt_stride=3))
    assert ck.n_checkpoints() == 2
    # null pointers are rejected before anything is launched
    rc = lib.dps_forward(plan.handle, 8, None, None, None, None, None, None, None, None, 0, None)
    assert lib.dps_forward_workspace_bytes(plan.handle, 8192) >= 4 * (8192 // 8 + 2)
    assert rc != 0 and b'null' in lib.dps_last_error()


def test_no_cpu_fallback():
    """Without a CUDA device the compute entry points refuse to run (the package has no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from tests import helpers
    sim, _, _ = helpers.sim_from_gold('ibb_heun', 'ibb', device='cpu')
    with pytest.raises(engine.EngineError, match='no CPU fallback'):
        sim.run()


def test_integration_doc_struct_matches_the_binding():
    """INTEGRATION.md shows the ctypes struct a reference maintainer would write; a stale copy (a missing trailing member)
    makes dps_plan_create read past the caller's struct.  Compare it field by field with engine.Topology and with the
    member order of include/dps.h."""
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    block = doc[doc.index('class dps_topology(ctypes.Structure)'):doc.index('DPS_ABI_VERSION = ')]
    doc_fields = []
    for name, ptr_to, plain in re.findall(r"\('(\w+)',\s*ctypes\.(?:POINTER\((?:ctypes\.)?(\w+)\)|(\w+))\)", block):
        doc_fields.append((name, 'ptr:' + ptr_to if ptr_to else plain))
    mine = []
    for name, ctype in engine.Topology._fields_:
        tname = ctype.__name__
        mine.append((name, 'ptr:' + tname[3:] if tname.startswith('LP_') else tname))
    def norm(t):             # ctypes.c_int32 is an alias of c_int on this ABI; compare the resolved type names
        pre, name = ('ptr:', t[4:]) if t.startswith('ptr:') else ('', t)
        name = {'dps_channel': 'Channel'}.get(name, name)
        return pre + (getattr(ctypes, name).__name__ if hasattr(ctypes, name) else name)
    assert [(n, norm(t)) for n, t in doc_fields] == [(n, norm(t)) for n, t in mine]
    # member order of the C header
    hdr = open(os.path.join(ROOT, 'include', 'dps.h')).read()
    body = hdr[hdr.index('typedef struct {\n    int32_t abi_version'):hdr.index('} dps_topology;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    assert re.findall(r'(\w+)\s*;', body) == [n for n, _ in mine]
    assert 'DPS_ABI_VERSION = %d' % engine.ABI_VERSION in doc


def test_prologue_entry_points_validate_their_arguments():
    """dps_load_flow / dps_dynamic_inverse / dps_model_init (initialize() on the device): argument errors are return codes
    with a message, before any CUDA call (so this runs without a GPU); rc 2 = outside the supported size."""
    import ctypes
    from diffpssi_b200 import prologue_device
    lib = engine.load_library()
    prologue_device._bind(lib)
    i32 = ctypes.c_int32
    types = (i32 * 70)(*([0] + [2] * 69))
    it = i32(0)
    assert lib.dps_load_flow(65, types, 1, None, 1, None, 1, None, 1, None, None, None, 1e-8, 10, ctypes.byref(it), None) == 2
    assert b'65 buses' in lib.dps_last_error()
    assert lib.dps_load_flow(3, types, 1, None, 1, None, 1, None, 1, None, None, None, 1e-8, 10, ctypes.byref(it), None) == 1
    assert lib.dps_dynamic_inverse(80, 1, None, 1, 0, None, None, 1, None, 1, None, None, None, None) == 2
    assert lib.dps_dynamic_inverse(5, 1, None, 1, 0, None, None, 1, None, 1, None, None, None, None) == 1
    zero = (i32 * 4)(0, 0, 0, 0)
    assert lib.dps_model_init(0, 5, 1, zero, zero, None, 1, None, 1, None, zero, None, None, None, None, None) == 2
    assert lib.dps_model_init(2, 5, 1, zero, zero, None, 1, None, 1, None, zero, None, None, None, None, None) == 1
    assert b'bad argument' in lib.dps_last_error()
    assert prologue_device.N_RAW == 35 and len(prologue_device.RAW) == prologue_device.N_RAW


def test_raw_parameter_order_matches_header():
    """enum dps_raw of include/dps.h and prologue_device.RAW (the host-side table builder) list the same parameters in the
    same order."""
    import os
    from diffpssi_b200 import prologue_device as pd
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'dps.h')).read()
    body = hdr[hdr.index('enum dps_raw {') + len('enum dps_raw {'):hdr.index('DPS_N_RAW\n};')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = [n.split('=')[0].strip() for n in body.split(',') if n.strip()]
    short = {'sn': 's_n', 'sn_sys': 's_n_sys', 'xd': 'x_d', 'xq': 'x_q', 'xdt': 'x_d_t', 'xqt': 'x_q_t', 'xdst': 'x_d_st',
             'xqst': 'x_q_st', 'td0t': 't_d0_t', 'tq0t': 't_q0_t', 'td0st': 't_d0_st', 'tq0st': 't_q0_st'}
    mine = [short.get(n[len('DPS_RAW_'):].lower(), n[len('DPS_RAW_'):].lower()) for n in names]
    assert mine == list(pd.RAW)
