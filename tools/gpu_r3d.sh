#!/bin/bash
# ncu --set full of the adjoint kernel (cm variant library)
mkdir -p gpurun_out/r3d
export DPS_B200_LIB=$PWD/gpurun_variants/libdps_${V:-cm}.so DPS_BWD_NO_PM3=1 CKPT=${CKPT:-8}
ACMD="python tools/adjoint_probe.py ieee39 8192"
$ACMD > gpurun_out/r3d/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dps_adjoint -s 1 -c 1 \
    -o gpurun_out/r3d/prof_adjoint_${V:-cm} $ACMD > gpurun_out/r3d/ncu_full_adj.log 2>&1
tail -2 gpurun_out/r3d/ncu_full_adj.log; tail -2 gpurun_out/r3d/plain.log
