#!/bin/bash
mkdir -p gpurun_out/r3f
O=gpurun_out/r3f
{
for V in soa nopf cm2; do
echo "== $V pm3 C=8"; DPS_B200_LIB=$PWD/gpurun_variants/libdps_$V.so CKPT=8 python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2 | head -1
echo "== $V pm1 C=8"; DPS_BWD_NO_PM3=1 DPS_B200_LIB=$PWD/gpurun_variants/libdps_$V.so CKPT=8 python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2 | head -1
done
} | tee $O/adjoint_ab.txt
