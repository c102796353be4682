"""``torch.autograd.Function`` around the stepping kernel.

Forward = ``dps_forward`` (one launch for the whole time loop).  Backward = ``dps_backward``, the hand-written
discrete adjoint over checkpointed steps: what the reference obtains by letting torch autograd record every op of
simulator.py:523-543 and calling ``loss.sum().backward()`` (optimization_lib/ps_optimization.py:212).
Gradients are returned with respect to the kernel inputs (x0, uniform constants, per-sim constants and -- when a network
parameter requires grad -- the reduced impedance matrices); torch then
continues through the prologue graph to the user's leaf tensors.
"""
import torch

from . import engine


class StepFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, plan, P, x0, upar, vpar, zsub):
        topo = plan.topo
        dev = x0.device
        n_rec, T, G = len(topo['rec']), topo['n_steps'], topo['n_gen']
        rec = torch.empty((n_rec, T, P), dtype=torch.float64, device=dev)
        x_final = torch.empty((G, engine.N_STATES, P), dtype=torch.float64, device=dev)
        n_ck = plan.n_checkpoints()
        ckpt = torch.empty((n_ck, G, engine.N_STATES, P), dtype=torch.float64, device=dev) if n_ck else None
        zsub = zsub.detach().contiguous()
        plan.forward(P, x0.detach().contiguous(), upar.detach().contiguous(),
                     None if vpar is None else vpar.detach().contiguous(), zsub, rec, x_final, ckpt)
        ctx.plan, ctx.P = plan, P
        ctx.has_vpar = vpar is not None
        ctx.save_for_backward(upar, vpar if vpar is not None else upar, zsub, ckpt if ckpt is not None else upar)
        ctx.has_ckpt = ckpt is not None
        ctx.rec_shape = tuple(rec.shape)
        # x_final is differentiable: a run that is cut into step ranges (parameter events) chains the final state of one
        # range into the initial state of the next, and the co-state travels back the same way (dps_backward: grad_x_final)
        ctx.set_materialize_grads(False)
        return rec, x_final

    @staticmethod
    def backward(ctx, g_rec, g_xf):
        from .adjoint import run_backward
        upar, vpar, zsub, ckpt = ctx.saved_tensors
        if not ctx.has_ckpt:
            raise engine.EngineError('backward needs checkpoints: run() saw no tensor requiring grad')
        if g_rec is None:
            g_rec = torch.zeros(ctx.rec_shape, dtype=torch.float64, device=upar.device)
        g_x0, g_upar, g_vpar, g_zsub = run_backward(ctx.plan, ctx.P, upar, vpar if ctx.has_vpar else None, zsub, ckpt,
                                                    g_rec.contiguous(), want_z=ctx.needs_input_grad[5],
                                                    g_xf=None if g_xf is None else g_xf.contiguous())
        return None, None, g_x0, g_upar, g_vpar, g_zsub
