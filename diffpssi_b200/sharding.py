"""Multi-GPU support: ``parallel_sims`` shards by contiguous rows, one process per GPU, nothing is exchanged in run().

Reference behaviour: the reference has no distributed code at all; every op of the time loop
is row-wise or a per-row matmul, so the batch axis shards with zero communication.  The only cross-shard quantities
live in the optimisation loop (optimization_lib/ps_optimization.py):

  * the per-sim loss vector ``[P]`` whose arg-min picks the best trajectory (ps_optimization.py:202)
    -> :func:`all_gather_rows`;
  * gradients of parameters that are *shared* by all sims (per-sim parameters need nothing: ``loss.sum().backward()``
    gives every row its own gradient, ps_optimization.py:212) -> :func:`all_reduce_shared_grads`.

Both are latency-bound messages of bytes to kilobytes; they go through ``torch.distributed`` (NCCL over NVLink on GPUs,
gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def shard_range(P_global, world_size=None, rank=None):
    """Rows [lo, hi) of the global batch owned by ``rank`` (contiguous, sizes differ by at most one)."""
    world_size = world_size if world_size is not None else get_world_size()
    rank = rank if rank is not None else get_rank()
    if P_global < world_size:
        # a rank with zero rows cannot build a simulation, would raise before the first collective and leave the other
        # ranks hanging in it: refuse on every rank, up front
        raise ValueError('parallel_sims=%d cannot be sharded over %d ranks (every rank needs at least one trajectory)'
                         % (P_global, world_size))
    return rank * P_global // world_size, (rank + 1) * P_global // world_size


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def init_distributed(backend=None):
    """One process per GPU, launched by torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* in the environment)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1 or dist.is_initialized():
        return
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda:%d' % local))
    else:
        dist.init_process_group(backend)


def shard_rows(array, P_global=None):
    """Slice the leading (batch) axis of a numpy array / tensor down to this rank's rows."""
    P_global = P_global if P_global is not None else array.shape[0]
    lo, hi = shard_range(P_global)
    return array[lo:hi]


def all_gather_rows(local, P_global):
    """Concatenate per-rank row blocks (e.g. the ``[P_local]`` loss vector) into the global ``[P_global, ...]`` tensor
    on every rank.  Shards may differ in size by one row, so blocks are padded to the largest shard for the collective."""
    world = get_world_size()
    if world == 1:
        return local
    sizes = [shard_range(P_global, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local.detach()
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def all_reduce_shared_grads(params):
    """Sum ``.grad`` of parameters shared by every sim across shards (one flat all-reduce)."""
    world = get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return
    flat = torch.cat([torch.view_as_real(g).reshape(-1) if g.is_complex() else g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for g in grads:
        n = g.numel() * (2 if g.is_complex() else 1)
        chunk = flat[off:off + n]
        if g.is_complex():
            g.copy_(torch.view_as_complex(chunk.reshape(g.shape + (2,))))
        else:
            g.copy_(chunk.reshape(g.shape))
        off += n
