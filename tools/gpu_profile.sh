#!/bin/bash
# usage: TAG=r1b [TESTS=1] bash tools/gpu_profile.sh   (run under gpurun; writes gpurun_out/*)
set -x
TAG=${TAG:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
if [ "${TESTS:-0}" = "1" ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_${TAG}.log
fi
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 1500 gpurun_out/bench_${TAG}.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
if [ "${LAUNCHES:-0}" = "1" ]; then
  $CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dps_ -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch.log 2>&1
fi
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dps_step_ -s 1 -c 1 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
