#!/bin/bash
# Final evidence of round 2 (run under gpurun, one GPU): bench line, ncu launch list of one bench run and of smoke(),
# ncu --set full of the stepping kernel of the headline run.
O=gpurun_out/final
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dps_ -c 40 --csv \
    --log-file $O/launches.csv $CMD > $O/ncu_launch.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file $O/smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > $O/ncu_smoke.log 2>&1
$CMD > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dps_step_ -s 4 -c 1 \
    -o $O/prof_step $CMD > $O/ncu_full_step.log 2>&1
tail -2 $O/ncu_full_step.log; tail -1 $O/smoke.log; head -c 400 $O/bench.json; echo
grep -o '"dps_[a-z_]*[^"]*"' $O/smoke_launches.csv | sort | uniq -c | head
