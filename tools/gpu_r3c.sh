#!/bin/bash
# adjoint: compile-time controller mask (CM 7) and constant-bank controller constants (PM 3)
mkdir -p gpurun_out/r3c
O=gpurun_out/r3c
{
echo "== cm7 + pm3"; DPS_B200_LIB=$PWD/gpurun_variants/libdps_cm.so CKPT=8 python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2
echo "== cm7 + pm1"; DPS_BWD_NO_PM3=1 DPS_B200_LIB=$PWD/gpurun_variants/libdps_cm.so CKPT=8 python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2
echo "== cm7 + pm3 stride 4"; DPS_B200_LIB=$PWD/gpurun_variants/libdps_cm.so CKPT=4 python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2
echo "== cm7 + pm3 stride 16"; DPS_B200_LIB=$PWD/gpurun_variants/libdps_cm.so CKPT=16 python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2
} | tee $O/adjoint_cm.txt
