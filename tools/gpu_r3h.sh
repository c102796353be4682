#!/bin/bash
# adjoint after the restructure (full library): gradient tests, strides, launch list
mkdir -p gpurun_out/r3h
O=gpurun_out/r3h
python -m pytest tests/test_cuda_gradients.py tests/test_cuda_optimization.py -x -q 2>&1 | tail -5 | tee $O/pytest_grad.log
{
for c in 4 8 16; do CKPT=$c python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2; done
python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2
for c in 1 8; do CKPT=$c python tools/adjoint_probe.py ibb 16384 2>&1 | tail -2; done
python tools/adjoint_probe.py ibb 16384 2>&1 | tail -2
for c in 1 8; do CKPT=$c python tools/adjoint_probe.py k2a 4096 2>&1 | tail -2; done
python tools/adjoint_probe.py k2a 16384 2>&1 | tail -2
python tools/adjoint_probe.py ieee39 16384 2>&1 | tail -2
} | tee $O/adjoint_probe.txt
python bench.py --config c4 --steps 10 2> $O/c4.err | tee $O/c4_n1.json | head -c 1200; echo
