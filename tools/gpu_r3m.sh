#!/bin/bash
# final adjoint kernel: small-system strides + ncu --set full (IEEE-39, 8192 sims, stride 8)
mkdir -p gpurun_out/r3m
O=gpurun_out/r3m
{
for c in 1 8; do CKPT=$c python tools/adjoint_probe.py ibb 16384 2>&1 | tail -2 | head -1; done
for c in 1 8; do CKPT=$c python tools/adjoint_probe.py k2a 16384 2>&1 | tail -2 | head -1; done
for c in 1 8; do CKPT=$c python tools/adjoint_probe.py k2a 4096 2>&1 | tail -2 | head -1; done
for c in 1 8; do CKPT=$c python tools/adjoint_probe.py ibb 2048 2>&1 | tail -2 | head -1; done
} | tee $O/adjoint_small.txt
ACMD="python tools/adjoint_probe.py ieee39 8192"
CKPT=8 ncu --set full --clock-control none --import-source on -k regex:dps_adjoint -s 1 -c 1 -o $O/prof_adjoint_final $ACMD > $O/ncu_full_adj.log 2>&1
tail -2 $O/ncu_full_adj.log
