#!/bin/bash
# Build an experimental variant of the engine that only carries the bench specialisation (G = 10):
#   tools/build_variant.sh <name> [extra nvcc flags, e.g. -DDPS_MIN_BLOCKS=4]
# -> gpurun_variants/libdps_<name>.so ; select it with DPS_B200_LIB=gpurun_variants/libdps_<name>.so
set -e
name=$1; shift
cd "$(dirname "$0")/../diffpssi_b200/csrc"
out=../../gpurun_variants; mkdir -p $out/obj_$name
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v -DDPS_ONLY_G10 $*"
for s in dps_abi dps_forward dps_forward_wide dps_network dps_backward dps_backward_wide dps_optim dps_prologue; do
  nvcc $F -c $s.cu -o $out/obj_$name/$s.o 2> $out/obj_$name/$s.log &
done
nvcc $F -DDPS_G=10 -c dps_forward_inst.cu -o $out/obj_$name/g10.o 2> $out/obj_$name/g10.log &
nvcc $F -DDPS_G=10 -c dps_backward_inst.cu -o $out/obj_$name/b10.o 2> $out/obj_$name/b10.log &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libdps_$name.so $out/obj_$name/*.o
grep -A2 "ILi10ELi7ELi3ELi0ELi1" $out/obj_$name/g10.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo
