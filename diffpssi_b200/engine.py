"""ctypes binding of the C ABI in ``include/dps.h`` (``libdps_b200.so``) -- the only route to the CUDA kernels.

There is deliberately **no CPU fallback**: if the shared library is missing, or no CUDA device is present,
every compute entry point raises.  PyTorch is used for device memory and streams only: tensors are passed as
raw device pointers, the current torch stream as ``cudaStream_t``.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DPS_B200_LIB') or os.path.join(_HERE, 'libdps_b200.so')   # override: kernel A/B experiments

ABI_VERSION = 4
N_STATES = 13
N_SLOTS = 42
N_FIELDS = 20
MAX_G = 16          # warp-exchange stepper
MAX_GW = 128        # wide stepper
MAX_EXTRA_OBS = 64
MAX_CFG = 64
MAX_FAULT = 16

EXPORTS = ('dps_plan_create', 'dps_plan_destroy', 'dps_n_checkpoints', 'dps_workspace_bytes', 'dps_forward_workspace_bytes',
           'dps_network_reduce', 'dps_forward', 'dps_forward_range', 'dps_backward', 'dps_fp64_peak_probe', 'dps_fp64_dmma_probe', 'dps_math_probe', 'dps_plan_info',
           'dps_last_error', 'dps_abi_version', 'dps_ident_loss_workspace_bytes', 'dps_ident_loss', 'dps_bfgs_step',
           'dps_ident_control', 'dps_load_flow', 'dps_dynamic_inverse', 'dps_model_init', 'dps_recorder_to_host')

SOLVER_ID = {'euler': 0, 'heun': 1, 'rk4': 2}
CH_GEN, CH_BUS_VRE, CH_BUS_VIM = 0, 1, 2


class EngineError(RuntimeError):
    pass


class Channel(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('index', ctypes.c_int32), ('field', ctypes.c_int32)]


class Topology(ctypes.Structure):
    _fields_ = [
        ('abi_version', ctypes.c_int32), ('n_gen', ctypes.c_int32), ('solver', ctypes.c_int32),
        ('n_steps', ctypes.c_int32), ('dt', ctypes.c_double), ('fn', ctypes.c_double),
        ('ctrl', ctypes.POINTER(ctypes.c_int32)),
        ('n_obs', ctypes.c_int32), ('z_rows', ctypes.c_int32), ('z_cols', ctypes.c_int32), ('n_cfg', ctypes.c_int32),
        ('step_cfg', ctypes.POINTER(ctypes.c_uint8)),
        ('z_per_sim', ctypes.c_int32),
        ('slot_vary', ctypes.POINTER(ctypes.c_int32)),
        ('n_vary', ctypes.c_int32), ('n_rec', ctypes.c_int32),
        ('rec', ctypes.POINTER(Channel)),
        ('ckpt_stride', ctypes.c_int32),
        ('common_ctrl', ctypes.POINTER(ctypes.c_double)),
    ]


_lib = None


def load_library(path=None):
    """dlopen the engine.  Raises EngineError when it has not been built (``python -c 'import __graft_entry__ as g;
    g.build()'`` or ``make -C diffpssi_b200/csrc``)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise EngineError('CUDA engine %s is missing: build it with `make -C diffpssi_b200/csrc` '
                          '(there is no CPU fallback by design)' % path)
    lib = ctypes.CDLL(path)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    lib.dps_last_error.restype = ctypes.c_char_p
    lib.dps_abi_version.restype = i32
    lib.dps_plan_create.argtypes = [ctypes.POINTER(Topology), ctypes.POINTER(vp)]
    lib.dps_plan_destroy.argtypes = [vp]
    lib.dps_n_checkpoints.argtypes = [vp]
    lib.dps_n_checkpoints.restype = i64
    lib.dps_recorder_to_host.argtypes = [vp, vp, i32, i64, i64, i64, vp]
    lib.dps_recorder_to_host.restype = i32
    lib.dps_workspace_bytes.argtypes = [vp, i64]
    lib.dps_workspace_bytes.restype = ctypes.c_size_t
    lib.dps_network_reduce.argtypes = [i64, i32, vp, vp, vp, vp, i32, ctypes.POINTER(i32), i32, ctypes.POINTER(i32),
                                       i32, ctypes.POINTER(i32), ctypes.POINTER(i32), dbl, vp, vp]
    lib.dps_forward_workspace_bytes.argtypes = [vp, i64]
    lib.dps_forward_workspace_bytes.restype = ctypes.c_size_t
    lib.dps_forward.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.dps_forward_range.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.dps_backward.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, ctypes.POINTER(ctypes.c_uint8), vp, vp, vp, vp,
                                 ctypes.c_size_t, vp]
    lib.dps_math_probe.argtypes = [i64, vp, vp, vp, vp]
    lib.dps_ident_loss_workspace_bytes.argtypes = [i64, i32]
    lib.dps_ident_loss_workspace_bytes.restype = ctypes.c_size_t
    lib.dps_ident_loss.argtypes = [i64, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.dps_bfgs_step.argtypes = [i64, i32, ctypes.POINTER(vp), ctypes.POINTER(i32), ctypes.POINTER(vp), ctypes.POINTER(i32),
                                  vp, vp, vp, vp, i32, i32, vp]
    lib.dps_ident_control.argtypes = [vp, dbl, vp]
    lib.dps_fp64_peak_probe.argtypes = [i32, ctypes.POINTER(dbl), ctypes.POINTER(dbl)]
    lib.dps_fp64_dmma_probe.argtypes = [i32, ctypes.POINTER(dbl), ctypes.POINTER(dbl)]
    lib.dps_plan_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32)]
    if lib.dps_abi_version() != ABI_VERSION:
        raise EngineError('libdps_b200.so has ABI %d, python side expects %d' % (lib.dps_abi_version(), ABI_VERSION))
    if path == LIB_PATH:
        _lib = lib
    return lib


def _check(lib, rc, what):
    if rc != 0:
        raise EngineError('%s failed (%d): %s' % (what, rc, lib.dps_last_error().decode()))


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise EngineError('engine buffers must live on a CUDA device (got %s); there is no CPU fallback' % t.device)
    if t.dtype != torch.float64 or not t.is_contiguous():
        raise EngineError('engine buffers must be contiguous float64')
    return ctypes.c_void_p(t.data_ptr())


def _stream(ref=None):
    """Current torch stream of the device that holds ``ref`` (default: the current device)."""
    dev = ref.device if ref is not None else None
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _on_device_of(t):
    """Context that makes the device of tensor ``t`` current for the duration of a C-ABI call: the library launches on
    the *current* device, so a simulation built with ``device='cuda:1'`` while device 0 is current would otherwise run
    its kernels on GPU 0 with GPU 1 pointers."""
    if t is None or not t.is_cuda:
        raise EngineError('engine buffers must live on a CUDA device; there is no CPU fallback')
    return torch.cuda.device(t.device)


def require_cuda():
    if not torch.cuda.is_available():
        raise EngineError("backend='cuda' needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")


class Plan(object):
    """Owns a ``dps_plan*``.  ``topology`` is a dict produced by ``planning.build_topology``."""

    def __init__(self, topo):
        self.lib = load_library()
        self.topo = topo
        G = topo['n_gen']
        self._ctrl = (ctypes.c_int32 * G)(*topo['ctrl'])
        T = topo['n_steps']
        self._step_cfg = (ctypes.c_uint8 * max(T, 1))(*topo['step_cfg'])
        self._slot_vary = (ctypes.c_int32 * (G * N_SLOTS))(*topo['slot_vary'])
        chans = topo['rec']
        self._rec = (Channel * max(len(chans), 1))(*[Channel(*c) for c in chans])
        t = Topology(ABI_VERSION, G, topo['solver'], T, topo['dt'], topo['fn'], self._ctrl,
                     topo['n_obs'], topo['z_rows'], topo['z_cols'], topo['n_cfg'], self._step_cfg,
                     topo['z_per_sim'], self._slot_vary, topo['n_vary'], len(chans), self._rec,
                     topo.get('ckpt_stride', 0), None)
        cc = topo.get('common_ctrl')
        if cc is not None:
            self._cc = (ctypes.c_double * N_SLOTS)(*cc)
            t.common_ctrl = self._cc
        self._t = t
        h = ctypes.c_void_p()
        _check(self.lib, self.lib.dps_plan_create(ctypes.byref(t), ctypes.byref(h)), 'dps_plan_create')
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.dps_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def n_checkpoints(self):
        return int(self.lib.dps_n_checkpoints(self.handle))

    def info(self):
        a, b, c, d = (ctypes.c_int32() for _ in range(4))
        _check(self.lib, self.lib.dps_plan_info(self.handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c),
                                                ctypes.byref(d)), 'dps_plan_info')
        return {'block_threads': a.value, 'sims_per_block': b.value, 'regs_per_thread': c.value,
                'g_template': d.value}

    def _forward_workspace(self, P, ref):
        """(pointer, bytes) of the optional forward scratch (work-queue counters), one cached buffer per device."""
        need = int(self.lib.dps_forward_workspace_bytes(self.handle, P))
        if need == 0:
            return None, 0
        ws = self.__dict__.setdefault('_fwd_ws', {}).get(ref.device)
        if ws is None or ws.numel() * 4 < need:
            ws = torch.empty((need + 3) // 4, dtype=torch.int32, device=ref.device)
            self._fwd_ws[ref.device] = ws
        return ctypes.c_void_p(ws.data_ptr()), ws.numel() * 4

    def forward(self, P, x0, upar, vpar, zsub, rec, x_final=None, ckpt=None):
        require_cuda()
        with _on_device_of(x0):
            ws, ws_bytes = self._forward_workspace(P, x0)
            _check(self.lib, self.lib.dps_forward(self.handle, P, _ptr(x0), _ptr(upar), _ptr(vpar), _ptr(zsub),
                                                  _ptr(rec), _ptr(x_final), _ptr(ckpt), ws, ws_bytes, _stream(x0)),
                   'dps_forward')

    def forward_range(self, P, step_begin, step_end, x_in, upar, vpar, zsub, rec, x_out, ckpt=None):
        require_cuda()
        with _on_device_of(x_in):
            ws, ws_bytes = self._forward_workspace(P, x_in)
            _check(self.lib, self.lib.dps_forward_range(self.handle, P, step_begin, step_end, _ptr(x_in), _ptr(upar),
                                                        _ptr(vpar), _ptr(zsub), _ptr(rec), _ptr(x_out), _ptr(ckpt),
                                                        ws, ws_bytes, _stream(x_in)), 'dps_forward_range')


def network_reduce(z_re, z_im, yd_re, yd_im, obs_bus, gen_bus, cfg_sets, y_fault, zsub):
    """On-device fault re-reduction (see include/dps.h).  ``z_*``: [n,n,Bz]; ``yd_*``: [n,Bz]; ``cfg_sets``: list of
    lists of faulted bus indices; ``zsub``: out [n_cfg, n_obs+m, n_gen+m, 2, Bz]."""
    require_cuda()
    lib = load_library()
    n_bus, Bz = z_re.shape[0], z_re.shape[2]
    ptr, flat = [0], []
    for s in cfg_sets:
        flat += list(s)
        ptr.append(len(flat))
    i32 = ctypes.c_int32
    a_obs = (i32 * len(obs_bus))(*obs_bus)
    a_gen = (i32 * len(gen_bus))(*gen_bus)
    a_ptr = (i32 * len(ptr))(*ptr)
    a_bus = (i32 * max(len(flat), 1))(*flat)
    with _on_device_of(zsub):
        _check(lib, lib.dps_network_reduce(Bz, n_bus, _ptr(z_re), _ptr(z_im), _ptr(yd_re), _ptr(yd_im),
                                           len(obs_bus), a_obs, len(gen_bus), a_gen, len(cfg_sets), a_ptr, a_bus,
                                           float(y_fault), _ptr(zsub), _stream(zsub)), 'dps_network_reduce')


def math_probe(a, b):
    """sin(a), cos(a), a/b, sqrt(|a|), 1/b computed by the kernels' own device primitives -> [5, n] tensor."""
    require_cuda()
    lib = load_library()
    out = torch.empty((5, a.numel()), dtype=torch.float64, device=a.device)
    with _on_device_of(a):
        _check(lib, lib.dps_math_probe(a.numel(), _ptr(a), _ptr(b), _ptr(out), _stream(a)), 'dps_math_probe')
    return out


def fp64_peak_probe(iters=4096):
    require_cuda()
    lib = load_library()
    tf, ms = ctypes.c_double(), ctypes.c_double()
    _check(lib, lib.dps_fp64_peak_probe(iters, ctypes.byref(tf), ctypes.byref(ms)), 'dps_fp64_peak_probe')
    return tf.value, ms.value


def fp64_dmma_probe(iters=4096):
    """FP64 tensor-core (DMMA) throughput of the current device in TFLOP/s -> (tflops, ms)."""
    require_cuda()
    lib = load_library()
    tf, ms = ctypes.c_double(), ctypes.c_double()
    _check(lib, lib.dps_fp64_dmma_probe(iters, ctypes.byref(tf), ctypes.byref(ms)), 'dps_fp64_dmma_probe')
    return tf.value, ms.value


# ----------------------------------------------------------------------------------------------------------------------
# identification step on the device (csrc/dps_optim.cu)
# ----------------------------------------------------------------------------------------------------------------------
def ident_loss(rec, target, vmin, vrange, state=None, want_grad=True, workspace=None):
    """Default loss of the reference's optimisation loop (ps_optimization.py:87-101, 178-199) on ``rec`` [C,T,P] against
    ``target`` [C,T,Pt] (Pt in {1, P}) -> (loss [P], g_rec [C,T,P] or None).  ``state``: [4] float64 device tensor that
    receives the batch minimum and its row (see include/dps.h)."""
    require_cuda()
    lib = load_library()
    C, T, P = rec.shape
    Pt = target.shape[2]
    loss = torch.empty(P, dtype=torch.float64, device=rec.device)
    g = torch.empty_like(rec) if want_grad else None
    need = int(lib.dps_ident_loss_workspace_bytes(P, T))
    if workspace is None or workspace.numel() * 8 < need:
        workspace = torch.empty((need + 7) // 8, dtype=torch.float64, device=rec.device)
    with _on_device_of(rec):
        _check(lib, lib.dps_ident_loss(P, T, C, Pt, _ptr(rec), _ptr(target), _ptr(vmin), _ptr(vrange), _ptr(loss), _ptr(g),
                                       _ptr(state), ctypes.c_void_p(workspace.data_ptr()), workspace.numel() * 8,
                                       _stream(rec)), 'dps_ident_loss')
    return loss, g, workspace


def _real_view(t):
    """(device pointer, element stride in doubles) of the real part of a [P,1] / [P] float64 or complex128 tensor."""
    if not t.is_cuda:
        raise EngineError('optimiser state must live on a CUDA device; there is no CPU fallback')
    if t.dtype == torch.complex128:
        r = torch.view_as_real(t)                     # [..., 2]
        return r.data_ptr(), 2 * (t.stride(0) if t.dim() else 1)
    if t.dtype != torch.float64:
        raise EngineError('optimisable parameters must be float64 or complex128')
    return t.data_ptr(), (t.stride(0) if t.dim() else 1)


def bfgs_step(params, grads, hesse, last_grad, last_sk, state, first_step, have_last):
    """CustomBFGSREALOptimizer.step() (optimizers.py:52-111) on the device; ``params`` are updated in place."""
    require_cuda()
    lib = load_library()
    n = len(params)
    P = params[0].shape[0]
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    xp, xs = zip(*[_real_view(p.data) for p in params])
    gp, gs = zip(*[_real_view(g) for g in grads])
    with _on_device_of(hesse):
        _check(lib, lib.dps_bfgs_step(P, n, (vp * n)(*xp), (i32 * n)(*xs), (vp * n)(*gp), (i32 * n)(*gs), _ptr(hesse),
                                      _ptr(last_grad), _ptr(last_sk), _ptr(state), int(first_step), int(have_last),
                                      _stream(hesse)), 'dps_bfgs_step')


def ident_control(state, decay):
    require_cuda()
    lib = load_library()
    with _on_device_of(state):
        _check(lib, lib.dps_ident_control(_ptr(state), float(decay), _stream(state)), 'dps_ident_control')
