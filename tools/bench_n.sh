#!/bin/bash
# bench.py on N GPUs the way the driver launches it (run under gpurun --gpus N): N=<n> TAG=<tag> bash tools/bench_n.sh
N=${N:-2}; TAG=${TAG:-n$N}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_${TAG}.json'))
print(d['n_gpus'], 'value %.4g' % d['value'], 'ms %.3f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'e2e %.4g' % d['e2e']['value'], d['kernel'], d['clocks']['reasons'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-200
