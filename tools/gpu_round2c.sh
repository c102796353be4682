#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SIMS="8192 65536" ROUNDS="0 16" EXTRA_ROUNDS="" bash tools/gpu_queue_ab.sh
for c in 32 16 8; do CKPT=$c python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2; done
for c in 32 8; do CKPT=$c python tools/adjoint_probe.py ibb 16384 2>&1 | tail -2; done
for c in 32 8; do CKPT=$c python tools/adjoint_probe.py k2a 4096 2>&1 | tail -2; done
