"""Host glue of the adjoint kernel: which constant slots need a gradient, the ``dps_backward`` call, and the scatter of
its per-trajectory slot gradients back onto the packed kernel inputs (upar / vpar)."""
import ctypes

import torch

from . import engine


def wanted_slots(init, slot_names):
    """uint8 mask [G*N_SLOTS]: 1 where the slot's source tensor is part of an autograd graph."""
    want = []
    for sl in init.slots:
        for name in slot_names:
            v = sl.get(name)
            want.append(1 if (v is not None and v.requires_grad) else 0)
    return want


def run_backward(plan, P, upar, vpar, zsub, ckpt, g_rec, want_z=False, g_xf=None):
    """-> (g_x0 [G,13,P], g_upar [G,N_SLOTS], g_vpar [n_vary,P] or None, g_zsub like zsub or None).
    ``g_xf`` [G,13,P]: cotangent of the final state (a later step range's g_x0), or None."""
    engine.require_cuda()
    lib = plan.lib
    topo = plan.topo
    G = topo['n_gen']
    dev = upar.device
    want = topo.get('want') or [0] * (G * engine.N_SLOTS)
    n_want = sum(want)
    # zero-filled: a run of zero steps writes neither (the library returns before launching anything)
    g_x0 = torch.zeros((G, engine.N_STATES, P), dtype=torch.float64, device=dev)
    gslot = torch.zeros((max(n_want, 1), P), dtype=torch.float64, device=dev)
    ws_bytes = int(lib.dps_workspace_bytes(plan.handle, P))
    ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device=dev)
    a_want = (ctypes.c_uint8 * len(want))(*want)
    # dL/dZsub is accumulated per trajectory ([n_cfg, rows, cols, 2, P], zero-initialised by the caller)
    gz = torch.zeros(tuple(zsub.shape[:4]) + (P,), dtype=torch.float64, device=dev) if want_z else None
    with engine._on_device_of(ckpt):
        rc = lib.dps_backward(plan.handle, P, engine._ptr(upar.detach().contiguous()),
                              engine._ptr(vpar.detach().contiguous()) if vpar is not None else None,
                              engine._ptr(zsub), engine._ptr(ckpt), engine._ptr(g_rec), engine._ptr(g_xf), a_want,
                              engine._ptr(g_x0), engine._ptr(gslot), engine._ptr(gz),
                              ctypes.c_void_p(ws.data_ptr()), ws_bytes, engine._stream(ckpt))
    engine._check(lib, rc, 'dps_backward')
    g_upar = torch.zeros_like(upar)
    g_vpar = torch.zeros_like(vpar) if vpar is not None else None
    slot_vary = topo['slot_vary']
    row = 0
    for i, w in enumerate(want):
        if not w:
            continue
        vi = slot_vary[i]
        if vi >= 0:
            g_vpar[vi] = gslot[row]
        else:
            g_upar[i // engine.N_SLOTS, i % engine.N_SLOTS] = gslot[row].sum()     # shared constant: sum over the batch
        row += 1
    if gz is not None and zsub.shape[-1] == 1:
        gz = gz.sum(dim=-1, keepdim=True)          # network shared by the batch: sum over the trajectories
    return g_x0, g_upar, g_vpar, gz
