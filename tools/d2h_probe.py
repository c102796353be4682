"""Where is the floor of the end-to-end arm at N GPUs?  Bare device->pinned-host copies of one recorder shard per rank,
all ranks at once and one rank alone, with the pinned buffer allocated (a) wherever the process happens to run and
(b) after binding the process to the CPUs of the GPU's own NUMA node (first touch then places the pages there).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/d2h_probe.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from diffpssi_b200 import hostmem


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda:%d' % local))
    n = 65536 // world * 2000 * 10                       # doubles of one rank's recorder shard
    dev = torch.empty(n, dtype=torch.float64, device='cuda').normal_()
    out = {'world': world, 'bytes_per_rank': n * 8}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(host, who):
        res = []
        for _ in range(4):
            barrier()
            t0 = time.perf_counter()
            if who is None or rank == who:
                host.copy_(dev, non_blocking=True)
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            res.append(dt)
        t = torch.tensor([min(res[1:])], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    info = hostmem.gpu_numa_info(local)
    for tag, bind in (('default', False), ('numa_local', True)):
        if bind:
            hostmem.bind_to_gpu_node(local)
        host = torch.empty(n, dtype=torch.float64).pin_memory()
        host.zero_()
        t_all = measure(host, None)
        t_solo = measure(host, 0)
        out[tag] = {'all_ranks_s': t_all, 'aggregate_GBps': world * n * 8 / t_all / 1e9,
                    'per_rank_GBps': n * 8 / t_all / 1e9, 'rank0_alone_GBps': n * 8 / t_solo / 1e9}
        del host
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    out['gpu_numa'] = infos
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
