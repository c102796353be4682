#!/bin/bash
# A/B of the work-queue scheduling (DPS_QUEUE_ROUNDS: 0 = static launch) for both stepping kernels at the shard sizes of
# the 1/2/4/8-GPU runs.  Run under gpurun; prints one line per run and keeps the JSON lines in gpurun_out/.
mkdir -p gpurun_out
run() {  # tag sims rounds pair_min
  DPS_QUEUE_ROUNDS=$3 DPS_PAIR_MIN_SIMS=$4 python bench.py --steps 10 --warmup 3 --sims $2 --no-cpu-baseline --no-e2e \
      > gpurun_out/q_$1.json 2> gpurun_out/q_$1.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/q_$1.json'))
    print('%-26s P=%-6d rounds=%-3s %8.3f ms  %.4g sim-steps/s  frac %.4f  regs %d spb %d checksum %.15g' % ('$1', $2, '$3', d['roofline']['kernel_ms'], d['value'], d['roofline']['frac'], d['kernel']['regs_per_thread'], d['kernel']['sims_per_block'], d['checksum']))
except Exception as e:
    print('$1', 'FAILED', e)
PY
}
for s in ${SIMS:-8192 16384 32768 65536}; do
  for r in ${ROUNDS:-0 16}; do
    run pair_${s}_r$r $s $r 0
    run single_${s}_r$r $s $r 999999999
  done
done
for r in ${EXTRA_ROUNDS:-6 10 24 40}; do
  run pair_8192_r$r 8192 $r 0
  run single_8192_r$r 8192 $r 999999999
done
