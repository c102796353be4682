"""Parameter identification on the device -- the reference's ``optimization_lib`` for ``backend='cuda'``.

Mirrors (same class names, constructor arguments, attribute names and printed lines):

    CustomBFGSREALOptimizer    src/diffpssi/optimization_lib/optimizers.py:7-111
    PowerSystemOptimization    src/diffpssi/optimization_lib/ps_optimization.py:17-256

What runs where.  One optimisation step of the reference is: ``sim.run()`` -> min/max normalisation -> loss ->
``loss.sum().backward()`` -> ``optimizer.step()`` -> step-size control (ps_optimization.py:170-249).  Here the simulation
and its adjoint are ``dps_forward`` / ``dps_backward`` (autograd.py); the default loss with its normalisation, the arg-min
over the batch and the loss gradient are ONE fused kernel pair (``dps_ident_loss``), the BFGS update is a per-trajectory
kernel (``dps_bfgs_step``) and the decay decision a one-thread kernel on a four-double device state
(``dps_ident_control``): no value crosses PCIe between two steps unless the caller asks for the printed progress lines.
With ``cuda_graph=True`` the whole step -- prologue, kernels, loss, adjoint, prologue backward, optimiser -- is captured
once into a CUDA graph and replayed, which removes the ~10 ms of host work per step that otherwise bounds small batches.

There is no CPU fallback: the kernels are the only implementation (``engine.py`` raises without the CUDA library).
"""
import os
import time

import torch

from . import engine


class _IdentLoss(torch.autograd.Function):
    """loss[p] = mean_t sum_c |target_norm - result_norm| (ps_optimization.py:87-101, 194-195) with its gradient with
    respect to the simulation result computed in the same pass (``dps_ident_loss``)."""

    @staticmethod
    def forward(ctx, result, target_ctp, vmin, vrange, state, holder):
        # result is [P,T,C]; the kernel wants the recorder's own memory order [C,T,P] (a free view for run() outputs)
        rec = result.detach().permute(2, 1, 0)
        if not rec.is_contiguous():
            rec = rec.contiguous()
        loss, g, holder['ws'] = engine.ident_loss(rec, target_ctp, vmin, vrange, state, want_grad=True,
                                                  workspace=holder.get('ws'))
        ctx.save_for_backward(g)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        (g,) = ctx.saved_tensors                       # [C,T,P]
        return (g * grad_loss.reshape(1, 1, -1)).permute(2, 1, 0), None, None, None, None, None


class CustomBFGSREALOptimizer(torch.optim.Optimizer):
    """BFGS without line search, first step scaled to the maximum relative step, every step clipped to it
    (optimizers.py:7-111).  State (inverse Hessian ``hesse`` [n,n,P], ``last_grads``, ``last_sk`` [n,P]) and the current
    maximum step live on the device; ``step()`` is one kernel launch."""

    def __init__(self, params, max_step=0.1, decay=0.9, verbose=True):
        params = list(params)
        if isinstance(max_step, (list, tuple)):
            raise NotImplementedError('per-parameter max_step lists are not supported on the device path')
        self.decay = decay
        self.verbose = verbose
        defaults = dict(max_steps=None)
        super(CustomBFGSREALOptimizer, self).__init__(params, defaults=defaults)
        P = params[0].shape[0]
        n = len(params)
        if any(p.shape[0] != P for p in params):
            raise ValueError('every optimisable parameter must be a [parallel_sims, 1] tensor (optimizers.py:33)')
        dev = params[0].device
        self.hesse = torch.zeros((n, n, P), dtype=torch.float64, device=dev)
        self.last_grads = torch.zeros((n, P), dtype=torch.float64, device=dev)
        self.last_sk = torch.zeros((n, P), dtype=torch.float64, device=dev)
        # {max_step, last_min_loss, cur_min_loss, cur_min_row}: read and written by the kernels (include/dps.h)
        self.state = torch.tensor([float(max_step), 0.0, 0.0, 0.0], dtype=torch.float64, device=dev)
        self.n_steps = 0

    @property
    def max_step(self):
        return float(self.state[0])                    # device -> host read; for printing only

    def decrease_step_size(self):
        """optimizers.py:40-50 (the identification loop leaves this decision to ``dps_ident_control``)."""
        self.state[0] *= self.decay
        if self.verbose:
            print('Decreasing max. step size to {}'.format(round(self.max_step, 4)))

    def step(self, **kwargs):
        """One BFGS step for every trajectory (one kernel launch).  Bookkeeping as optimizers.py:64-110: the first call
        initialises the inverse Hessian from the gradients and only stores (grads, s); from the second call on the Hessian is
        updated with the stored pair."""
        params = self.param_groups[0]['params']
        grads = [p.grad for p in params]
        if any(g is None for g in grads):
            raise RuntimeError('step() before backward(): a parameter has no gradient')
        engine.bfgs_step(params, grads, self.hesse, self.last_grads, self.last_sk, self.state,
                         first_step=(self.n_steps == 0), have_last=(self.n_steps > 0))
        self.n_steps += 1


optimizer_dict = {'bfgs': CustomBFGSREALOptimizer}      # ps_optimization.py:12-14


class PowerSystemOptimization(object):
    """ps_optimization.py:17-256 for a ``diffpssi_b200`` simulation.  Same constructor; ``run(max_steps)`` prints the same
    progress lines (set ``verbose=False`` to keep the loop free of device->host reads).

    Extensions: ``verbose``; ``cuda_graph`` (capture one step into a CUDA graph and replay it; needs the default loss and a
    network that does not change between steps); ``check_every`` (with ``verbose=False``: how often the loss threshold is
    tested on the host)."""

    def __init__(self, sim, original_data, params_optimizable, param_names=None, optimizer='bfgs', params_original=None,
                 max_step=0.1, decay=0.9, enable_plots=False, normalize_loss=True, loss_function=None,
                 loss_threshold=None, verbose=True, cuda_graph=False, check_every=10):
        self.sim = sim
        if getattr(sim, 'backend', None) != 'cuda':
            raise NotImplementedError("diffpssi_b200.PowerSystemOptimization drives backend='cuda' simulations only")
        if enable_plots:
            raise NotImplementedError('plots are outside the accelerated path (matplotlib is not a dependency)')
        self.params = list(params_optimizable)
        dev = self.params[0].device
        self.target_data = original_data.to(dev)
        self.optimizer = optimizer_dict[optimizer](self.params, max_step=max_step, decay=decay, verbose=verbose)
        self.params_original = params_original
        self.last_min_loss = None
        self.loss_threshold = loss_threshold if loss_threshold else 0
        self.param_names = param_names or ['Param {}'.format(i) for i in range(len(self.params))]
        self.normalize_loss = normalize_loss
        self.loss_function = loss_function                # None: the fused default loss
        self.verbose = verbose
        self.cuda_graph = cuda_graph
        self.check_every = max(int(check_every), 1)
        self.decay = decay
        self._graph = None
        self._holder = {}
        tgt = self.target_data
        if tgt.dim() == 2:
            tgt = tgt.unsqueeze(0)
        self._target3 = tgt                               # [Pt,T,C]
        if normalize_loss:
            # ps_optimization.py:103-113
            self.min_values = torch.min(tgt, dim=1)[0].unsqueeze(1)
            self.max_values = torch.max(tgt, dim=1)[0].unsqueeze(1)
            self.range_values = self.max_values - self.min_values
            if bool(torch.any(self.range_values == 0)):
                if verbose:
                    print('Using simulation data to normalize loss function')
            elif verbose:
                print('Using target data to normalize loss function')
        self._pack_target()

    # ---- device-side views of target / normalisation in the kernel's [C,T,Pt] layout ----------------------------------
    def _pack_target(self):
        tgt = self._target3
        uniform = tgt.shape[0] == 1 or bool((tgt == tgt[:1]).all())
        t = tgt[:1] if uniform else tgt
        self._target_ctp = t.permute(2, 1, 0).contiguous().to(torch.float64)
        if self.normalize_loss:
            lo, rg = self.min_values, self.range_values
            if uniform:
                lo, rg = lo[:1], rg[:1]
            self._vmin = lo[:, 0, :].permute(1, 0).contiguous().to(torch.float64)       # [C,Pt]
            self._vrange = rg[:, 0, :].permute(1, 0).contiguous().to(torch.float64)
        else:
            self._vmin = self._vrange = None

    def _normalisation_from_results(self, results):
        """ps_optimization.py:180-190: target range is degenerate -> median over the batch of the simulated min / max."""
        min_values = torch.min(results, dim=1)[0].unsqueeze(1)
        max_values = torch.max(results, dim=1)[0].unsqueeze(1)
        self.min_values = torch.median(min_values, dim=0)[0].unsqueeze(0).detach()
        self.max_values = torch.median(max_values, dim=0)[0].unsqueeze(0).detach()
        self.range_values = self.max_values - self.min_values
        self._pack_target()

    # ---- one step ---------------------------------------------------------------------------------------------------------
    def _loss(self, results):
        if self.loss_function is not None:
            # a user-supplied torch loss: normalise as the reference does and let autograd differentiate it
            if self.normalize_loss:
                target_norm = (self._target3 - self.min_values) / self.range_values
                res_norm = (results - self.min_values) / self.range_values
            else:
                target_norm, res_norm = self._target3, results
            loss = self.loss_function(res_norm, target_norm)
            vals = torch.nan_to_num(loss.detach(), 100000)
            m, i = vals.min(dim=0)
            self.optimizer.state[2] = m
            self.optimizer.state[3] = i.to(torch.float64)
            return loss
        return _IdentLoss.apply(results, self._target_ctp, self._vmin, self._vrange, self.optimizer.state, self._holder)

    def _step_body(self, k):
        """simulate -> loss -> backward -> optimiser -> step-size control; everything stays on the device."""
        for p in self.params:
            p.grad = None
        t, results = self.sim.run()
        if self.normalize_loss and k == 0 and bool(torch.any(self.range_values == 0)):
            self._normalisation_from_results(results.detach())
        loss = self._loss(results)
        loss.sum().backward()
        return t, results, loss

    def _optimizer_step(self, k):
        self.optimizer.step()                                        # ps_optimization.py:235
        self.sim.reset()                                             # :242
        engine.ident_control(self.optimizer.state, self.decay)       # :244-249

    def _capture(self, k):
        """Capture step ``k`` (k >= 1: steady-state optimiser flags) into a CUDA graph.  Everything the host computes while
        building the step (event schedule, plan, which slots vary ...) is frozen into the graph, which is valid because
        it depends on the grid and on which tensors are leaves, not on their values; the memoised prologue parts (load
        flow, inv(Y)) must not depend on a leaf."""
        graph = torch.cuda.CUDAGraph()
        for p in self.params:
            p.grad = None
        with torch.cuda.graph(graph, stream=self._stream):
            t, results, loss = self._step_body(k)
            self._optimizer_step(k)
        self._graph = (graph, t, results, loss)

    def run(self, max_steps=100):
        if os.environ.get('DIFFPSSI_FORCE_OPT_ITERS') is not None:
            max_steps = int(os.environ.get('DIFFPSSI_FORCE_OPT_ITERS'))
            print('WARNING: FORCING THE USE OF {} OPTIMIZATION ITERATION.'
                  'THIS SHOULD ONLY HAPPEN FOR UNITTESTS'.format(os.environ.get('DIFFPSSI_FORCE_OPT_ITERS')))
        graphed = self.cuda_graph and self.loss_function is None
        if not graphed:
            return self._run(max_steps, False)
        # The eager warm-up steps, the capture and the replays all run on ONE side stream: autograd binds a leaf's
        # gradient accumulation to the stream of its first backward, and a capture that has to synchronise with another
        # stream is invalid.
        if getattr(self, '_stream', None) is None:
            self._stream = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        self._stream.wait_stream(cur)
        try:
            with torch.cuda.stream(self._stream):
                return self._run(max_steps, True)
        finally:
            cur.wait_stream(self._stream)

    def _run(self, max_steps, graphed):
        opt_start_time = time.time()
        t = results = loss = None
        self.history = []
        for opt_step in range(max_steps):
            if graphed and opt_step >= 3:
                if self._graph is None:
                    self._capture(opt_step)            # capturing does not execute: the replay below is this step
                graph, t, results, loss = self._graph
                graph.replay()
                stop = self._report(opt_step, loss, after_step=True)
            else:
                t, results, loss = self._step_body(opt_step)
                stop = self._report(opt_step, loss, after_step=False)
                if not stop:
                    self._optimizer_step(opt_step)
            if stop:
                break
        if self.verbose:
            print('Optimization finished in {:.2f} seconds'.format(time.time() - opt_start_time))
        self.results, self.time = results, t
        return t, results

    def _report(self, opt_step, loss, after_step):
        """Printed progress (ps_optimization.py:204-232) and the loss-threshold test.  Reads four doubles and the best row's
        parameters back; with ``verbose=False`` only every ``check_every`` steps (and only if a threshold is set)."""
        state = self.optimizer.state
        if not self.verbose:
            if self.loss_threshold and (opt_step + 1) % self.check_every == 0:
                return float(state[2]) < self.loss_threshold
            return False
        host = state.tolist()
        min_loss_val, min_loss_idx = host[2], int(host[3])
        self.history.append((min_loss_idx, min_loss_val))
        print('Step: {}, Min. Loss Batch: {}, Min. Loss: {}'.format(opt_step, min_loss_idx, min_loss_val))
        best = [float(p.detach()[min_loss_idx].real) for p in self.params]
        note = ' (after the update of this step)' if after_step else ''
        print('Current Best Params{}: '.format(note) +
              ', '.join(['{}: {:.3f}'.format(self.param_names[i], best[i]) for i in range(len(best))]))
        if self.params_original is not None:
            print('Relative Errors in Percent: ' +
                  ', '.join(['{}: {:.2f}%'.format(self.param_names[i], (best[i] - self.params_original[i]) /
                                                  self.params_original[i] * 100) for i in range(len(best))]))
        print('-' * 100)
        if min_loss_val < self.loss_threshold:
            print('Loss threshold reached. Optimization stopped.')
            return True
        return False


class GraphedStep(object):
    """Run ``fn()`` -- typically ``sim.reset(); sim.run(); loss(...).sum().backward()`` -- through one captured CUDA graph.

    The first ``warmup`` calls execute eagerly (they fill the prologue's memoisation caches and the allocator), the next
    call captures, every later call replays: ~10 us of host work per step instead of the ~10 ms of Python / autograd
    bookkeeping that bound small batches.  ``fn`` must be free of device->host reads, its leaf tensors must be updated in
    place (``p.data.copy_``) and everything it returns lives at fixed addresses (read it before the next call).
    All calls run on one private stream (see PowerSystemOptimization.run); the caller's current stream waits for it."""

    def __init__(self, fn, warmup=3):
        self.fn, self.warmup = fn, warmup
        self.calls = 0
        self.graph = None
        self.out = None
        self.stream = torch.cuda.Stream()

    def eager(self):
        """``fn()`` without the graph, on the same private stream.  (Autograd binds the gradient accumulation of a leaf to
        the stream of the first graph that used it and keeps that binding while anything -- e.g. the simulation's cached
        prologue results -- references that graph; an eager step on another stream before the capture would make the
        capture depend on that stream, which CUDA rejects.  So: every execution of ``fn`` goes through this object.)"""
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            out = self.fn()
        cur.wait_stream(self.stream)
        return out

    def __call__(self):
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            if self.calls < self.warmup:
                self.out = self.fn()
            else:
                if self.graph is None:
                    self.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph, stream=self.stream):
                        self.out = self.fn()
                self.graph.replay()
        self.calls += 1
        cur.wait_stream(self.stream)
        return self.out
