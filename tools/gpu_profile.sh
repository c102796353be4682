set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG:-r1a}.json 2> gpurun_out/bench_${TAG:-r1a}.err
tail -c 3000 gpurun_out/bench_${TAG:-r1a}.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG:-r1a}.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -5 gpurun_out/ncu_launch.log
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dps_step_kernel -s 1 -c 1 -o gpurun_out/prof_${TAG:-r1a} $CMD > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
