"""BASELINE config 4 on N GPUs: one parameter-identification step on the IBB model -- forward + adjoint gradients of the
trajectory MSE w.r.t. per-sim H, D, X_d' (16384 sims sharded by rows) plus one parameter *shared* by all sims, whose
gradient is all-reduced over NCCL; the per-sim loss vector is all-gathered (arg-min of ps_optimization.py:202).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ident_step.py

Rank 0 also runs the whole batch on its own GPU and checks that the sharded result reproduces it."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from diffpssi_b200 import Pss, sharding
from tests import helpers

P_GLOBAL, T = 16384, 2000


def build(rows, shared_value, params, target):
    P = rows[1] - rows[0]
    sim = Pss(parallel_sims=P, sim_time=T * 0.005, time_step=0.005, solver='heun', grid_data=helpers.load_grid('ibb'),
              verbose=False)
    sim.add_sc_event(1, 1.05, 'Bus 1')
    leaves = helpers.apply_overrides(sim, {k: v[rows[0]:rows[1]] for k, v in params.items()}, requires_grad=True)
    shared = torch.tensor([[shared_value + 0j]], dtype=torch.complex128, device='cuda', requires_grad=True)
    sim.get_generator_by_name('Gen 1').t_q0_st = shared
    sim.set_record_channels(['Gen 1.omega', 'Gen 1.e_q_st', 'Gen 1.e_d_st'])

    def step():
        for l in list(leaves.values()) + [shared]:
            l.grad = None
        sim.reset()
        _, res = sim.run()
        loss = ((res - target[None]) ** 2).mean(dim=(1, 2))
        loss.sum().backward()
        return loss.detach()
    return step, leaves, shared


def main():
    sharding.init_distributed()
    world, rank = sharding.get_world_size(), sharding.get_rank()
    rng = np.random.RandomState(0)
    params = {('Gen 1', 'h'): 3.5 * rng.uniform(0.5, 2, P_GLOBAL), ('Gen 1', 'd'): rng.uniform(0, 2, P_GLOBAL),
              ('Gen 1', 'x_d_t'): 0.3 * rng.uniform(0.5, 2, P_GLOBAL)}
    target = torch.from_numpy(np.load(os.path.join(helpers.GOLD, 'ref_ibb_model.npy'))).cuda()
    lo, hi = sharding.shard_range(P_GLOBAL)
    step, leaves, shared = build((lo, hi), 0.07, params, target)

    def full_step():
        loss = step()
        full = sharding.all_gather_rows(loss, P_GLOBAL)
        sharding.all_reduce_shared_grads([shared])
        return full
    for _ in range(2):
        full_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        full = full_step()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / n], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    out = {'config': 'C4 IBB identification step, heun, P=%d T=%d, forward + adjoint + collectives' % (P_GLOBAL, T),
           'n_gpus': world, 'ms_per_step': float(dt) * 1e3, 'sim_steps_per_s': P_GLOBAL * T / float(dt),
           'best_row': int(torch.argmin(torch.nan_to_num(full, nan=float('inf'))))}
    if rank == 0:
        ref_step, ref_leaves, ref_shared = build((0, P_GLOBAL), 0.07, params, target)
        ref_loss = ref_step()
        ok = torch.isfinite(ref_loss)
        out['loss_max_abs_diff'] = float((full - ref_loss)[ok].abs().max())
        g = {k: v.grad.real.reshape(-1) for k, v in leaves.items()}
        gr = {k: v.grad.real.reshape(-1)[lo:hi] for k, v in ref_leaves.items()}
        out['per_sim_grad_max_rel_diff'] = max(float(((g[k] - gr[k]).abs() / (gr[k].abs() + 1e-300))[torch.isfinite(gr[k])].max()) for k in g)
        out['shared_grad'] = float(shared.grad.real)
        out['shared_grad_reference'] = float(ref_shared.grad.real)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
