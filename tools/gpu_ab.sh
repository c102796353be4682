#!/bin/bash
# A/B of kernel variants on the GPU box (run under gpurun).
#   TAG=<tag> [TESTS=0|1] VARIANTS="name|ENV=.. ENV=..;name2|..." [PROFILE="name name2"] bash tools/gpu_ab.sh
set -x
TAG=${TAG:-ab}
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_${TAG}.log
fi
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 900 gpurun_out/bench_${TAG}.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
IFS=';' read -ra VS <<< "$VARIANTS"
for v in "${VS[@]}"; do
  name=${v%%|*}; envs=${v#*|}
  env $envs python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_${name}.json 2> gpurun_out/bench_${TAG}_${name}.err
  tail -c 700 gpurun_out/bench_${TAG}_${name}.json
done
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dps_step_kernel -s 1 -c 1 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full.log 2>&1
for v in "${VS[@]}"; do
  name=${v%%|*}; envs=${v#*|}
  if [[ " $PROFILE " == *" $name "* ]]; then
    env $envs $CMD > gpurun_out/plain3.log 2>&1 && env $envs ncu --set full --clock-control none --import-source on -k regex:dps_step_kernel -s 1 -c 1 -o gpurun_out/prof_${TAG}_${name} $CMD > gpurun_out/ncu_full_${name}.log 2>&1
  fi
done
