#!/bin/bash
# adjoint occupancy A/B: register caps 168 / 128 (libraries from tools/build_variant.sh)
mkdir -p gpurun_out/r3b
O=gpurun_out/r3b
{
for v in bwd3 bwd4; do
  for c in 8 32; do
    echo "== $v"; DPS_B200_LIB=$PWD/gpurun_variants/libdps_$v.so CKPT=$c python tools/adjoint_probe.py ieee39 8192 2>&1 | tail -2
  done
done
} | tee $O/adjoint_variants.txt
