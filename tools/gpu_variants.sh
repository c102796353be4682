#!/bin/bash
# A/B of engine variants built by tools/build_variant.sh (run under gpurun).
#   TAG=<tag> VARIANTS="a b c" [SIMS="65536 8192"] [PROFILE="a"] bash tools/gpu_variants.sh
TAG=${TAG:-ab}
mkdir -p gpurun_out
for v in $VARIANTS; do
  for s in ${SIMS:-65536}; do
    DPS_B200_LIB=$PWD/gpurun_variants/libdps_$v.so python bench.py --steps 5 --warmup 3 --sims $s --no-cpu-baseline --no-e2e \
      > gpurun_out/bench_${TAG}_${v}_$s.json 2> gpurun_out/bench_${TAG}_${v}_$s.err
    python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_${TAG}_${v}_$s.json'))
    print('%-14s P=%-6d %8.3f ms  %.4g sim-steps/s  frac %.4f  regs %d  checksum %.15g' % ('$v', $s, d['roofline']['kernel_ms'], d['value'], d['roofline']['frac'], d['kernel']['regs_per_thread'], d['checksum']))
except Exception as e:
    print('$v', $s, 'FAILED', e)
PY
  done
done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
for v in $TRAFFIC; do      # light capture: DRAM bytes of one stepping-kernel launch
  DPS_B200_LIB=$PWD/gpurun_variants/libdps_$v.so ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active --clock-control none -k regex:dps_step_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_${TAG}_$v.csv $CMD > /dev/null 2>&1
  echo "$v: $(grep -o '"\(dram__bytes\|smsp__inst\|sm__inst\)[^,]*","[^"]*","[^"]*"' gpurun_out/traffic_${TAG}_$v.csv | tr '\n' ' ')"
done
for v in $PROFILE; do
  DPS_B200_LIB=$PWD/gpurun_variants/libdps_$v.so ncu --set full --clock-control none --import-source on -k regex:dps_step_kernel -s 1 -c 1 -o gpurun_out/prof_${TAG}_$v $CMD > gpurun_out/ncu_full_$v.log 2>&1
done
