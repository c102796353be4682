#!/bin/bash
# ncu capture of the wide stepper on BASELINE config 5 (run under gpurun): TAG=<tag> bash tools/prof_c5.sh
TAG=${TAG:-c5}
python tools/bench_configs.py c5 > gpurun_out/bench_c5_$TAG.json 2>&1
cat gpurun_out/bench_c5_$TAG.json
ncu --set full --clock-control none --import-source on -k regex:dps_step_wide_kernel -s 1 -c 1 -o gpurun_out/prof_c5_$TAG python tools/bench_configs.py c5 > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log
