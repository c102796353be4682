#!/bin/bash
# round 3 (second half of round 2), first call: adjoint with the fat segment scratch
mkdir -p gpurun_out/r3a
O=gpurun_out/r3a
python -m pytest tests/test_cuda_gradients.py -x -q 2>&1 | tail -5 > $O/pytest_grad.log
cat $O/pytest_grad.log
{
for c in 32 16 8 4; do CKPT=$c python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2; done
for c in 32 8 1; do CKPT=$c python tools/adjoint_probe.py ibb 16384 2>&1 | tail -2; done
for c in 32 8 1; do CKPT=$c python tools/adjoint_probe.py k2a 4096 2>&1 | tail -2; done
CKPT=8 python tools/adjoint_probe.py ieee39 16384 2>&1 | tail -2
} | tee $O/adjoint_probe.txt
