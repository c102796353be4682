#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_cuda_gradients.py -x -q 2>&1 | tail -3
for c in 32 1; do CKPT=$c python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2; done
for c in 32 1; do CKPT=$c python tools/adjoint_probe.py ibb 16384 2>&1 | tail -2; done
for c in 32 1; do CKPT=$c python tools/adjoint_probe.py ibb 2048 2>&1 | tail -2; done
for c in 32 1; do CKPT=$c python tools/adjoint_probe.py k2a 4096 2>&1 | tail -2; done
python bench.py --config c4 --steps 10 2> gpurun_out/c4.err | tee gpurun_out/c4_n1.json | head -c 900; echo
python bench.py --config c4 --steps 10 --sims 2048 2>> gpurun_out/c4.err | tee gpurun_out/c4_sims2048.json | head -c 900; echo
