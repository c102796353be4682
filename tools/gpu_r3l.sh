#!/bin/bash
# adjoint sweep with cp.async staging of the stage records (full library)
mkdir -p gpurun_out/r3l
O=gpurun_out/r3l
python -m pytest tests/test_cuda_gradients.py tests/test_cuda_optimization.py -x -q 2>&1 | tail -3 | tee $O/pytest_grad.log
{
for c in 4 8; do CKPT=$c python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2 | head -1; done
CKPT=8 python tools/adjoint_probe.py k2a 4096 2>&1 | tail -2 | head -1
python tools/adjoint_probe.py k2a 16384 2>&1 | tail -2 | head -1
python tools/adjoint_probe.py ibb 16384 2>&1 | tail -2 | head -1
python tools/adjoint_probe.py ieee39 16384 2>&1 | tail -2 | head -1
} | tee $O/adjoint_async.txt
