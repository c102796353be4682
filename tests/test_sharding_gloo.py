"""world_size-2 gloo tests (CPU) of the N>1 path: row sharding of parallel_sims, the loss all-gather and the all-reduce of
shared-parameter gradients (diffpssi_b200/sharding.py).  The kernels are replaced by their host emulation
(tests/helpers.py); everything else -- prologue, planning, autograd chaining, collectives -- is the package's own code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers

P_GLOBAL = 5        # odd on purpose: shards of 2 and 3 rows


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _params(P):
    rng = np.random.RandomState(7)
    return {('Gen 1', 'h'): 3.5 * rng.uniform(0.8, 1.5, P), ('Gen 1', 'd'): rng.uniform(0.5, 2.0, P)}


def _run(rows, shared_value):
    """Forward + backward on the given global rows; returns (loss [rows], grads of per-sim leaves, grad of the shared leaf)."""
    from diffpssi_b200 import Pss
    P = rows[1] - rows[0]
    sim = Pss(parallel_sims=P, sim_time=0.6, time_step=0.005, solver='heun', grid_data=helpers.load_grid('ibb'),
              verbose=False, device='cpu')
    sim.add_sc_event(0.1, 0.15, 'Bus 1')
    ov = {k: v[rows[0]:rows[1]] for k, v in _params(P_GLOBAL).items()}
    leaves = helpers.apply_overrides(sim, ov, requires_grad=True)
    shared = torch.tensor([[shared_value + 0j]], dtype=torch.complex128, requires_grad=True)
    sim.get_generator_by_name('Gen 1').t_q0_st = shared
    sim.set_record_channels(['Gen 1.omega', 'Gen 1.e_q_st'])
    result = helpers.run_emulated_autograd(sim)
    loss = (result ** 2).mean(dim=(1, 2))
    loss.sum().backward()
    return loss.detach(), {k: v.grad.real.reshape(-1).clone() for k, v in leaves.items()}, shared


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from diffpssi_b200 import sharding
    sharding.init_distributed('gloo')
    lo, hi = sharding.shard_range(P_GLOBAL)
    loss, grads, shared = _run((lo, hi), 0.07)
    full_loss = sharding.all_gather_rows(loss, P_GLOBAL)
    sharding.all_reduce_shared_grads([shared])
    out[rank] = dict(range=(lo, hi), loss=full_loss.numpy(), grads={k: v.numpy() for k, v in grads.items()},
                     shared_grad=shared.grad.real.item())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    ref_loss, ref_grads, ref_shared = _run((0, P_GLOBAL), 0.07)
    assert out[0]['range'] == (0, 2) and out[1]['range'] == (2, 5)
    for r in range(world):
        assert np.allclose(out[r]['loss'], ref_loss.numpy(), rtol=1e-12, atol=0)            # every rank has the full vector
        lo, hi = out[r]['range']
        for k, g in out[r]['grads'].items():
            assert np.allclose(g, ref_grads[k].numpy()[lo:hi], rtol=1e-10, atol=0)             # per-sim grads stay local
        assert np.isclose(out[r]['shared_grad'], ref_shared.grad.real.item(), rtol=1e-10)      # shared grad summed over shards


def test_shard_ranges_cover_the_batch():
    from diffpssi_b200 import sharding
    for P, world in ((65536, 8), (7, 3), (8, 8)):
        cuts = [sharding.shard_range(P, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == P
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        sizes = [hi - lo for lo, hi in cuts]
        assert max(sizes) - min(sizes) <= 1


def test_fewer_trajectories_than_ranks_is_rejected_up_front():
    """A rank with zero rows would raise before the first collective and leave the others hanging in it."""
    import pytest
    from diffpssi_b200 import sharding
    with pytest.raises(ValueError, match='cannot be sharded'):
        sharding.shard_range(5, 8, 0)
