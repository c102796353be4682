#!/bin/bash
# A/B of dps_step_wide_kernel variants (tools/build_variant.sh ... -DDPS_WIDE_*) on BASELINE config 5
mkdir -p gpurun_out
for v in $VARIANTS; do
  DPS_B200_LIB=$PWD/gpurun_variants/libdps_$v.so timeout 300 python bench.py --config c5 --steps 2 --warmup 3 > gpurun_out/c5_$v.json 2> gpurun_out/c5_$v.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/c5_$v.json'))
    print('%-10s %8.2f ms  %.4g sim-steps/s  %.2f TFLOP/s  frac %.4f  finite %.3f' % ('$v', d['ms_per_step'], d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['finite_fraction']))
except Exception as e:
    print('$v FAILED', e)
PY
done
