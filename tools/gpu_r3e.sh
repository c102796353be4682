#!/bin/bash
mkdir -p gpurun_out/r3e
O=gpurun_out/r3e
V=${V:-cm2}
{
echo "== $V pm3 C=8"; DPS_B200_LIB=$PWD/gpurun_variants/libdps_$V.so CKPT=8 python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2
echo "== $V pm1 C=8"; DPS_BWD_NO_PM3=1 DPS_B200_LIB=$PWD/gpurun_variants/libdps_$V.so CKPT=8 python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2
echo "== $V pm3 C=4"; DPS_B200_LIB=$PWD/gpurun_variants/libdps_$V.so CKPT=4 python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2
echo "== $V pm3 C=16"; DPS_B200_LIB=$PWD/gpurun_variants/libdps_$V.so CKPT=16 python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2
} | tee $O/adjoint_$V.txt
