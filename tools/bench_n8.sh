#!/bin/bash
# 8-GPU confirmation: the headline bench (config 3) and the identification step (config 4), launched like the driver does.
N=${N:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "c3 rc=$?"; head -c 700 gpurun_out/bench_n$N.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config c4 --steps 20 --warmup 5 \
    > gpurun_out/c4_n$N.json 2> gpurun_out/c4_n$N.err
echo "c4 rc=$?"; head -c 1500 gpurun_out/c4_n$N.json; echo; tail -3 gpurun_out/c4_n$N.err
