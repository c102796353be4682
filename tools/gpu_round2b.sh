#!/bin/bash
# The gpurun command sequence of the second half of round 2 (adjoint restructure, initialize() on the device):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2b.sh'
# Variant A/Bs of the adjoint (controller mask, scratch layout, register caps, cp.async staging) were built with
#   tools/build_variant.sh <name> -DDPS_BWD_AOS=1 | -DDPS_BWD_L1PF=0 | -DDPS_BWD_ASYNC=0 | -DDPS_BWD_MIN_BLOCKS=3 | -DDPS_BWD_WARPS=3 -DDPS_BWD_MAXNREG=224
# and timed with DPS_B200_LIB=gpurun_variants/libdps_<name>.so CKPT=<stride> python tools/adjoint_probe.py ieee39 8192.
O=gpurun_out/round2b
mkdir -p $O
python -m pytest tests/test_cuda_gradients.py tests/test_cuda_optimization.py tests/test_cuda_prologue.py -x -q 2>&1 | tail -3 | tee $O/pytest.log
{
for c in 4 8 16; do CKPT=$c python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2 | head -1; done
python tools/adjoint_probe.py ieee39 16384 2>&1 | tail -2 | head -1
for c in 1 8; do CKPT=$c python tools/adjoint_probe.py k2a 4096 2>&1 | tail -2 | head -1; done
for c in 1 8; do CKPT=$c python tools/adjoint_probe.py ibb 16384 2>&1 | tail -2 | head -1; done
} | tee $O/adjoint_probe.txt
python tools/prologue_probe.py 2>&1 | tee $O/prologue_probe.txt
ACMD="python tools/adjoint_probe.py ieee39 8192"
CKPT=8 ncu --set full --clock-control none --import-source on -k regex:dps_adjoint -s 1 -c 1 -o $O/prof_adjoint $ACMD > $O/ncu_full_adj.log 2>&1
tail -2 $O/ncu_full_adj.log
