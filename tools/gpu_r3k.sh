#!/bin/bash
# wide stepper: dense packing (416-thread blocks, 152 registers) vs padded (512 threads, 128 registers) on config 5
mkdir -p gpurun_out/r3k
O=gpurun_out/r3k
export DPS_B200_LIB=$PWD/gpurun_variants/libdps_${V:-dense}.so
python -m pytest tests/test_cuda_parity.py -q -m gpu -k "wide or tiled or config5" 2>&1 | tail -3 | tee $O/pytest_wide.log
python bench.py --config c5 --steps 3 --warmup 3 > $O/c5_dense.json 2> $O/c5_dense.err; python -c "
import json; d=json.load(open('$O/c5_dense.json')); print('dense ', d['ms_per_step'], d['roofline'])"
DPS_WIDE_PADDED=1 python bench.py --config c5 --steps 3 --warmup 3 > $O/c5_padded.json 2> $O/c5_padded.err; python -c "
import json; d=json.load(open('$O/c5_padded.json')); print('padded', d['ms_per_step'], d['roofline'])"
