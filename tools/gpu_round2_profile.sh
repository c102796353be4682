#!/bin/bash
# Round-2 evidence (run under gpurun, one GPU): bench line, ncu launch list of one bench run, ncu --set full of the stepping
# kernel of the headline run (work-queue build of dps_step_pair_kernel) and of the adjoint kernel (IEEE-39, 8192 sims).
mkdir -p gpurun_out
TAG=${TAG:-round2}
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dps_ -c 40 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dps_step_ -s 4 -c 1 \
    -o gpurun_out/prof_${TAG}_step $CMD > gpurun_out/ncu_full_step.log 2>&1
tail -2 gpurun_out/ncu_full_step.log
ACMD="python tools/adjoint_probe.py ieee39 8192"
CKPT=32 $ACMD > gpurun_out/plain3.log 2>&1 && CKPT=32 ncu --set full --clock-control none --import-source on -k regex:dps_adjoint -s 1 -c 1 \
    -o gpurun_out/prof_${TAG}_adjoint $ACMD > gpurun_out/ncu_full_adj.log 2>&1
tail -2 gpurun_out/ncu_full_adj.log
ls -la gpurun_out/*.ncu-rep
